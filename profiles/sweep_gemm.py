"""Sweep the grouped GEMM launch time over problem count / K / split-K (tuning aid, GPU only)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sac_td3_td7_b200 import _lib  # noqa: E402

lib = _lib.lib()
torch.cuda.set_device(0)
dev = torch.device("cuda")
B, H = 256, 256


def launch(mode, nprob, K, splitk, M=B):
    os.environ["B200RL_SPLITK"] = str(splitk)
    keep, probs = [], []
    for _ in range(nprob):
        p = _lib.GemmProblem()
        if mode == 0:    # Y[M,H] = X[M,K] W[H,K]^T
            x, w, b, y = torch.randn(M, K, device=dev), torch.randn(H, K, device=dev), torch.randn(H, device=dev), torch.empty(M, H, device=dev)
            p.I, p.J, p.C, p.c_ld, p.act, p.bias = M, H, y.data_ptr(), H, 2, b.data_ptr()
            s = p.seg[0]; s.A, s.a_ld, s.B, s.b_ld, s.len = x.data_ptr(), K, w.data_ptr(), K, K
        elif mode == 1:  # dX[M,H] = dZ[M,K] W[K,H]
            x, w, y = torch.randn(M, K, device=dev), torch.randn(K, H, device=dev), torch.empty(M, H, device=dev)
            b = None
            p.I, p.J, p.C, p.c_ld = M, H, y.data_ptr(), H
            s = p.seg[0]; s.A, s.a_ld, s.B, s.b_ld, s.len = x.data_ptr(), K, w.data_ptr(), H, K
        else:            # dW[H,K] = dZ[M,H]^T X[M,K]
            x, w, y = torch.randn(M, H, device=dev), torch.randn(M, K, device=dev), torch.empty(H, K, device=dev)
            b = None
            p.I, p.J, p.C, p.c_ld = H, K, y.data_ptr(), K
            s = p.seg[0]; s.A, s.a_ld, s.B, s.b_ld, s.len = x.data_ptr(), H, w.data_ptr(), K, M
        p.nseg = 1
        keep += [x, w, b, y]
        probs.append(p)
    arr = (_lib.GemmProblem * nprob)(*probs)
    ms = C.c_float()
    _lib.check(lib.b200rl_op_gemm_timed(mode, arr, nprob, 1, 20, 300, C.byref(ms), torch.cuda.current_stream().cuda_stream))
    return ms.value * 1e3


print("mode nprob K splitk -> us per launch")
for mode in (0, 1, 2):
    for nprob in (1, 2, 5):
        for K in (256, 768):
            row = []
            for sk in (1, 2, 3, 4):
                row.append(f"{launch(mode, nprob, K, sk):6.2f}")
            print(mode, nprob, K, " ".join(row), flush=True)
# empty-ish kernel cost reference: K=64 single chunk
print("K=64:", " ".join(f"{launch(0, n, 64, 1):6.2f}" for n in (1, 2, 5)))
