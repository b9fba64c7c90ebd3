"""Forward ops of ``rl.nn`` on CUDA tensors through the C ABI (no autograd: inference only).

``linear`` consumes a list of input tensors as the column-wise concatenation the reference builds
with ``torch.cat`` (sale.py:52,80,114-117) without materialising it.
"""
from __future__ import annotations

import torch

from ... import _lib

ACT = {None: 0, "none": 0, "relu": 1, "elu": 2, "tanh": 3}


_stream = _lib.current_stream_ptr


def _chk(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.B200RLError("sac_td3_td7_b200.rl.nn forwards run on CUDA tensors only (no CPU fallback)")
    return t.detach().float().contiguous()


def linear(xs, weight: torch.Tensor, bias: torch.Tensor, act: str | None = None) -> torch.Tensor:
    """act(cat(xs, 1) @ weight.T + bias) -- one fused GEMM launch."""
    lib = _lib.lib()
    xs = [_chk(x) for x in (xs if isinstance(xs, (list, tuple)) else [xs])]
    w, b = _chk(weight), _chk(bias)
    M, N, K = xs[0].shape[0], w.shape[0], w.shape[1]
    assert sum(x.shape[1] for x in xs) == K and len(xs) <= 3
    y = torch.empty((M, N), device=w.device, dtype=torch.float32)
    p = _lib.GemmProblem()
    p.I, p.J, p.C, p.c_ld, p.act = M, N, y.data_ptr(), N, ACT[act]
    p.bias = b.data_ptr()
    off = 0
    for x in xs:
        s = p.seg[p.nseg]
        s.A, s.a_ld, s.B, s.b_ld, s.len = x.data_ptr(), x.shape[1], w.data_ptr() + 4 * off, K, x.shape[1]
        p.nseg += 1
        off += x.shape[1]
    arr = (_lib.GemmProblem * 1)(p)
    _lib.check(lib.b200rl_op_gemm(0, arr, 1, 1, _stream()))
    return y


def avg_l1_norm(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """sale.py:11-13 (eps is fixed at 1e-8 in the kernel, the only value the reference uses)."""
    assert eps == 1e-8
    x = _chk(x)
    y = torch.empty_like(x)
    _lib.check(_lib.lib().b200rl_op_avgl1norm(0, x.data_ptr(), None, y.data_ptr(), x.shape[0], x.shape[1], _stream()))
    return y
