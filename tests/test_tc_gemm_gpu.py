"""tcgen05 (TF32) grouped GEMM vs torch fp32 -- the stacked-seed tensor-core path, forced on for
single-group launches with B200RL_TC=1.  Tolerance: TF32 inputs (10-bit mantissa), fp32 accumulate:
<= 4e-3 normwise (stated TF32 tolerance of the north star; the single-agent path stays fp32)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from sac_td3_td7_b200 import _lib
from test_op_kernels import ACTS, _problem, _run, _seg, nerr

pytestmark = pytest.mark.gpu
TOL = 4e-3  # operands are truncated to TF32 (2^-10 relative per operand), fp32 accumulate


@pytest.fixture()
def tc(dev, monkeypatch):
    if dev.backend != "cuda":
        pytest.skip("tensor-core path exists only in the CUDA build")
    monkeypatch.setenv("B200RL_TC", "1")
    return dev


@pytest.mark.parametrize("M,N,Ks,act", [(256, 256, (256,), 2), (128, 256, (64,), 0), (256, 256, (256, 256, 256), 1), (256, 256, (256, 6), 2),
                                        (256, 128, (256,), 0), (256, 64, (40,), 3), (512, 256, (17,), 2)])
def test_tc_forward(tc, M, N, Ks, act):
    dev = tc
    rng = np.random.default_rng(M + N)
    K = sum(Ks)
    ld = (K + 3) // 4 * 4
    xs = [rng.standard_normal((M, k)).astype(np.float32) for k in Ks]
    W = np.zeros((N, ld), np.float32)
    W[:, :K] = rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)
    b = rng.standard_normal(N).astype(np.float32)
    y = dev.zeros((M, N))
    Wd = dev.to_dev(W)
    p = _problem(M, N, y, N)
    off = 0
    for x, k in zip(xs, Ks):
        _seg(p, dev.to_dev(x), k, Wd + 4 * off, ld, k)
        off += k
    p.bias, p.act = dev.to_dev(b), act
    _run(dev, 0, [p])
    ref = ACTS[act](F.linear(torch.tensor(np.concatenate(xs, 1)), torch.tensor(W[:, :K]), torch.tensor(b))).numpy()
    assert nerr(dev.from_dev(y, (M, N)), ref) < TOL


@pytest.mark.parametrize("act,ncols,k_lo", [(0, 256, 0), (2, 256, 0), (1, 256, 256), (3, 64, 16)])
def test_tc_dgrad(tc, act, ncols, k_lo):
    dev = tc
    rng = np.random.default_rng(act + ncols)
    M, N, K = 256, 256, 528
    dZ = rng.standard_normal((M, N)).astype(np.float32)
    W = rng.standard_normal((N, K)).astype(np.float32)
    pre = rng.standard_normal((M, ncols)).astype(np.float32)
    H = ACTS[act](torch.tensor(pre)).numpy()
    dX = dev.zeros((M, ncols))
    p = _problem(M, ncols, dX, ncols)
    _seg(p, dev.to_dev(dZ), N, dev.to_dev(W) + 4 * k_lo, K, N)
    if act:
        p.aux, p.aux_ld, p.act = dev.to_dev(H), ncols, act
    _run(dev, 1, [p])
    xt = torch.tensor(pre, requires_grad=True)
    ACTS[act](xt).backward(torch.tensor(dZ) @ torch.tensor(W[:, k_lo:k_lo + ncols]))
    assert nerr(dev.from_dev(dX, (M, ncols)), xt.grad.numpy()) < TOL


def test_tc_dgrad_two_segment_sum(tc):
    dev = tc
    rng = np.random.default_rng(1)
    M, N, Z = 256, 256, 256
    dz = [rng.standard_normal((M, N)).astype(np.float32) for _ in range(2)]
    Ws = [rng.standard_normal((N, 768)).astype(np.float32) for _ in range(2)]
    dX = dev.zeros((M, Z))
    p = _problem(M, Z, dX, Z)
    for d, w in zip(dz, Ws):
        _seg(p, dev.to_dev(d), N, dev.to_dev(w) + 4 * 256, 768, N)
    _run(dev, 1, [p])
    ref = sum(torch.tensor(d) @ torch.tensor(w[:, 256:512]) for d, w in zip(dz, Ws)).numpy()
    assert nerr(dev.from_dev(dX, (M, Z)), ref) < TOL


def test_tc_wgrad_with_fused_adam(tc):
    dev = tc
    rng = np.random.default_rng(2)
    M, N, K = 256, 256, 256
    dZ = rng.standard_normal((M, N)).astype(np.float32) / M
    x = rng.standard_normal((M, K)).astype(np.float32)
    W = rng.standard_normal((N, K)).astype(np.float32) * 0.05
    p_d, m_d, v_d, g_d = dev.to_dev(W), dev.zeros(W.shape), dev.zeros(W.shape), dev.zeros(W.shape)
    sc = _lib.AdamScalars(w1=1 - 0.9, beta2=0.999, w2=1 - 0.999, bc2_sqrt=math.sqrt(1 - 0.999), neg_step=-(3e-4 / (1 - 0.9)), eps=1e-8, step=1)
    q = _problem(N, K, g_d, K)
    _seg(q, dev.to_dev(dZ), N, dev.to_dev(x), K, M)
    q.P, q.M, q.V, q.p_ld, q.adam = p_d, m_d, v_d, K, dev.to_dev(np.frombuffer(bytes(sc), np.uint8))
    _run(dev, 2, [q])
    gref = (torch.tensor(dZ).T @ torch.tensor(x)).numpy()
    g = dev.from_dev(g_d, W.shape)
    assert nerr(g, gref) < TOL
    # Adam from the kernel's own gradient must be exact arithmetic on that gradient
    wt = torch.tensor(W.copy(), requires_grad=True)
    opt = torch.optim.Adam([wt], lr=3e-4)
    wt.grad = torch.tensor(g)
    opt.step()
    assert np.abs(dev.from_dev(p_d, W.shape) - wt.detach().numpy()).max() <= 2e-7
    assert np.array_equal(dev.from_dev(m_d, W.shape), opt.state[wt]["exp_avg"].numpy())


def test_tc_stacked_groups_and_mixed_table(tc):
    """groups > 1 with an eligible (tensor-core) and an ineligible (FFMA, N = 6) problem in one table."""
    dev = tc
    rng = np.random.default_rng(3)
    G, M, N, K, A = 3, 256, 256, 256, 6
    x = rng.standard_normal((G, M, K)).astype(np.float32)
    W = rng.standard_normal((G, N, K)).astype(np.float32) / 16
    Wa = rng.standard_normal((G, A, K)).astype(np.float32) / 16
    y, ya = dev.zeros((G, M, N)), dev.zeros((G, M, A))
    xd = dev.to_dev(x)
    p = _problem(M, N, y, N)
    _seg(p, xd, K, dev.to_dev(W), K, K)
    p.seg[0].a_gs, p.seg[0].b_gs, p.c_gs = M * K, N * K, M * N
    pa = _problem(M, A, ya, A)
    _seg(pa, xd, K, dev.to_dev(Wa), K, K)
    pa.seg[0].a_gs, pa.seg[0].b_gs, pa.c_gs, pa.act = M * K, A * K, M * A, 3
    _run(dev, 0, [p, pa], groups=G)
    ref = torch.einsum("gmk,gnk->gmn", torch.tensor(x), torch.tensor(W)).numpy()
    refa = torch.tanh(torch.einsum("gmk,gnk->gmn", torch.tensor(x), torch.tensor(Wa))).numpy()
    assert nerr(dev.from_dev(y, (G, M, N)), ref) < TOL
    assert nerr(dev.from_dev(ya, (G, M, A)), refa) < 2e-6  # fp32 FFMA path


@pytest.mark.parametrize("env", [{"B200RL_TC_WS": "0"}, {"B200RL_TC_TMA": "0"}])
def test_tc_kernel_variants_in_a_fresh_process(env):
    """The kernel variant is chosen once per process (persistent warp-specialised + TMA by default): run this module's
    other tests again with the one-tile-per-CTA TMA kernel, and with the cp.async staging, each in its own process."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_tc_gemm_gpu.py"), "-q", "-x", "-m", "gpu", "-k", "not fresh_process",
                        "-p", "no:cacheprovider"], env=e, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("rows,out,dz_ld", [(256, 256, 256), (256, 256, 768), (130, 64, 64), (256, 6, 6), (67, 20, 24), (256, 1, 1)])
def test_bias_gradient_column_sums(tc, rows, out, dz_ld):
    """bias_grad_kernel (bias gradients of the tensor-core path): db = column sums of dZ as four chains over r mod 4
    combined (s0 + s1) + (s2 + s3) -- the 128-bit path (out % 4 == 0, aligned) and the scalar path give that same fp32
    sum bit for bit -- then fused Adam; stacked seeds."""
    dev = tc
    G = 3
    rng = np.random.default_rng(rows + out)
    dZ = rng.standard_normal((G, rows, dz_ld)).astype(np.float32)
    b0 = rng.standard_normal((G, out)).astype(np.float32)
    p_d, m_d, v_d, g_d = dev.to_dev(b0), dev.zeros(b0.shape), dev.zeros(b0.shape), dev.zeros(b0.shape)
    sc = _lib.AdamScalars(w1=1 - 0.9, beta2=0.999, w2=1 - 0.999, bc2_sqrt=math.sqrt(1 - 0.999), neg_step=-(3e-4 / (1 - 0.9)), eps=1e-8, step=1)
    q = _problem(out, 1, g_d, 1)
    _seg(q, dev.to_dev(dZ), dz_ld, None, 0, rows, ones=1)
    q.seg[0].a_gs, q.c_gs, q.p_gs = rows * dz_ld, out, out
    q.P, q.M, q.V, q.p_ld, q.adam = p_d, m_d, v_d, 1, dev.to_dev(np.frombuffer(bytes(sc), np.uint8))
    _run(dev, 2, [q], groups=G)
    d = dZ[:, :, :out]
    r4 = rows & ~3
    chains = [d[:, k:r4:4, :] for k in range(4)]
    s = []
    for k in range(4):
        acc = np.zeros((G, out), np.float32)
        for r in range(chains[k].shape[1]):
            acc = acc + chains[k][:, r, :]
        s.append(acc)
    for r in range(r4, rows):
        s[0] = s[0] + d[:, r, :]
    want = (s[0] + s[1]) + (s[2] + s[3])
    g = dev.from_dev(g_d, b0.shape)
    assert np.array_equal(g, want)
    wt = torch.tensor(b0.copy(), requires_grad=True)
    opt = torch.optim.Adam([wt], lr=3e-4)
    wt.grad = torch.tensor(g)
    opt.step()
    assert np.abs(dev.from_dev(p_d, b0.shape) - wt.detach().numpy()).max() <= 2e-7



FP32_TOL = 2e-6


def _adam_scalars(dev, t=1, lr=3e-4):
    sc = _lib.AdamScalars(w1=1 - 0.9, beta2=0.999, w2=1 - 0.999, bc2_sqrt=math.sqrt(1 - 0.999 ** t), neg_step=-(lr / (1 - 0.9 ** t)), eps=1e-8, step=t)
    return dev.to_dev(np.frombuffer(bytes(sc), np.uint8))


def test_wgrad_with_a_non_tileable_input_width_is_cut_into_column_blocks(tc):
    """376 input columns (Humanoid) = tcgen05 blocks of 256 and 96 columns (TF32 tolerance) + a 24-column narrow remainder (fp32)."""
    dev = tc
    rng = np.random.default_rng(12)
    G, M, N, K = 8, 256, 256, 376
    dZ = rng.standard_normal((G, M, N)).astype(np.float32) / M
    x = rng.standard_normal((G, M, K)).astype(np.float32)
    W = (rng.standard_normal((G, N, K)) * 0.05).astype(np.float32)
    p_d, m_d, v_d, g_d = dev.to_dev(W), dev.zeros(W.shape), dev.zeros(W.shape), dev.zeros(W.shape)
    q = _problem(N, K, g_d, K)
    _seg(q, dev.to_dev(dZ), N, dev.to_dev(x), K, M)
    q.seg[0].a_gs, q.seg[0].b_gs, q.c_gs, q.p_gs = M * N, M * K, N * K, N * K
    q.P, q.M, q.V, q.p_ld, q.adam = p_d, m_d, v_d, K, _adam_scalars(dev)
    _run(dev, 2, [q], groups=G)
    g, out = dev.from_dev(g_d, W.shape), dev.from_dev(p_d, W.shape)
    for s in (0, G - 1):
        ref = (torch.tensor(dZ[s]).T @ torch.tensor(x[s])).numpy()
        assert nerr(g[s][:, :352], ref[:, :352]) < TOL
        assert nerr(g[s][:, 352:], ref[:, 352:]) < FP32_TOL
        wt = torch.tensor(W[s].copy(), requires_grad=True)
        opt = torch.optim.Adam([wt], lr=3e-4)
        wt.grad = torch.tensor(g[s])
        opt.step()
        assert np.abs(out[s] - wt.detach().numpy()).max() <= 2e-7


@pytest.mark.parametrize("G,M,N,K", [(3, 256, 256, 256), (2, 512, 256, 27), (1, 128, 64, 40)])
def test_tc_forward_with_fused_avgl1norm(tc, G, M, N, K):
    """act = 5 (kActNorm): C = x W^T + b and aux = AvgL1Norm(C) = C / max(mean|C|, 1e-8) per row (rl/nn/sale.py:11-13), the
    normalisation done in the epilogue of the producing layer by the warps that share a row group."""
    dev = tc
    rng = np.random.default_rng(G * 1000 + K)
    ld = (K + 3) // 4 * 4
    x = np.zeros((G, M, ld), np.float32)
    x[:, :, :K] = rng.standard_normal((G, M, K))
    W = np.zeros((G, N, ld), np.float32)
    W[:, :, :K] = rng.standard_normal((G, N, K)) / np.sqrt(K)
    b = rng.standard_normal((G, N)).astype(np.float32)
    y, yn = dev.zeros((G, M, N)), dev.zeros((G, M, N))
    p = _problem(M, N, y, N)
    _seg(p, dev.to_dev(x), ld, dev.to_dev(W), ld, K)
    p.seg[0].a_gs, p.seg[0].b_gs, p.c_gs = M * ld, N * ld, M * N
    p.bias, p.bias_gs, p.act = dev.to_dev(b), N, 5
    p.aux, p.aux_ld, p.aux_gs = yn, N, M * N
    _run(dev, 0, [p], groups=G)
    got, gotn = dev.from_dev(y, (G, M, N)), dev.from_dev(yn, (G, M, N))
    for s in range(G):
        ref = F.linear(torch.tensor(x[s, :, :K]), torch.tensor(W[s, :, :K]), torch.tensor(b[s]))
        refn = ref / ref.abs().mean(-1, keepdim=True).clamp(min=1e-8)
        assert nerr(got[s], ref.numpy()) < TOL
        assert nerr(gotn[s], refn.numpy()) < TOL
        # the normalised output is exactly the kernel's own pre-norm output over its row mean
        own = got[s] / np.maximum(np.abs(got[s]).mean(-1, keepdims=True), 1e-8)
        assert nerr(gotn[s], own) < 2e-6
