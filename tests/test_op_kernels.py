"""Op-level kernels through the C ABI vs a plain torch fp32 reference of the same op.

Tolerances (fp32): GEMM-family outputs <= 2e-6 normwise (max|d| / max|ref|) -- fp32 FMA
accumulation in a different order than MKL; AvgL1Norm / SAC sample <= 2e-6; Adam given
identical inputs <= 1e-6 normwise on p and exact-order on m, v; Polyak bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from sac_td3_td7_b200 import _lib


pytestmark = pytest.mark.usefixtures("gemm_path")  # both product loops of the small-tile kernel (conftest.py)


def nerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _problem(I, J, Cptr, c_ld):
    p = _lib.GemmProblem()
    p.I, p.J, p.C, p.c_ld = I, J, Cptr, c_ld
    return p


def _seg(p, A, a_ld, B, b_ld, ln, ones=0):
    s = p.seg[p.nseg]
    s.A, s.B, s.a_ld, s.b_ld, s.len, s.b_ones = A, B, a_ld, b_ld, ln, ones
    p.nseg += 1


def _run(dev, mode, probs, groups=1):
    arr = (_lib.GemmProblem * len(probs))(*probs)
    dev.check(dev.lib.b200rl_op_gemm(mode, arr, len(probs), groups, None))


ACTS = {0: lambda x: x, 1: F.relu, 2: F.elu, 3: torch.tanh}


@pytest.mark.parametrize("M,N,Ks,act", [(256, 256, (256,), 2), (64, 256, (17, 6), 0), (100, 40, (256, 256, 256), 1), (256, 1, (256,), 0),
                                        (33, 6, (250,), 3), (1, 256, (376,), 2)])
def test_linear_forward_with_concat_segments(dev, M, N, Ks, act):
    rng = np.random.default_rng(M * 7 + N)
    K = sum(Ks)
    ld = (K + 3) // 4 * 4
    xs = [rng.standard_normal((M, k)).astype(np.float32) for k in Ks]
    W = np.zeros((N, ld), np.float32)
    W[:, :K] = rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)
    b = rng.standard_normal(N).astype(np.float32)
    y = dev.zeros((M, N))
    Wd, bd = dev.to_dev(W), dev.to_dev(b)
    p = _problem(M, N, y, N)
    off = 0
    for x, k in zip(xs, Ks):
        _seg(p, dev.to_dev(x), k, Wd + 4 * off, ld, k)
        off += k
    p.bias, p.act = bd, act
    _run(dev, 0, [p])
    ref = ACTS[act](F.linear(torch.tensor(np.concatenate(xs, 1)), torch.tensor(W[:, :K]), torch.tensor(b))).numpy()
    assert nerr(dev.from_dev(y, (M, N)), ref) < 2e-6


def test_forward_target_smoothing_epilogue(dev):
    rng = np.random.default_rng(0)
    M, N, K = 256, 6, 256
    x = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / 16).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    noise = rng.standard_normal((M, N)).astype(np.float32) * 3
    y = dev.zeros((M, N))
    p = _problem(M, N, y, N)
    _seg(p, dev.to_dev(x), K, dev.to_dev(W), K, K)
    p.bias, p.act, p.aux, p.aux_ld, p.noise_sigma, p.noise_clip = dev.to_dev(b), 4, dev.to_dev(noise), N, 0.2, 0.5
    _run(dev, 0, [p])
    t = torch.tanh(F.linear(torch.tensor(x), torch.tensor(W), torch.tensor(b)))
    ref = (t + (torch.tensor(noise) * 0.2).clamp(-0.5, 0.5)).clamp(-1, 1).numpy()
    assert np.abs(dev.from_dev(y, (M, N)) - ref).max() < 2e-6


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_dgrad_with_activation_derivative_and_column_window(dev, act):
    rng = np.random.default_rng(act)
    M, N, K, k_lo, ncols = 256, 256, 279, 17, 6 if act == 3 else 256
    dZ = rng.standard_normal((M, N)).astype(np.float32)
    ld = (K + 3) // 4 * 4
    W = np.zeros((N, ld), np.float32)
    W[:, :K] = rng.standard_normal((N, K)).astype(np.float32)
    pre = rng.standard_normal((M, ncols)).astype(np.float32)
    H = ACTS[act](torch.tensor(pre)).numpy()
    dX = dev.zeros((M, ncols))
    p = _problem(M, ncols, dX, ncols)
    _seg(p, dev.to_dev(dZ), N, dev.to_dev(W) + 4 * k_lo, ld, N)
    if act:
        p.aux, p.aux_ld, p.act = dev.to_dev(H), ncols, act
    _run(dev, 1, [p])
    xt = torch.tensor(pre, requires_grad=True)
    up = torch.tensor(dZ) @ torch.tensor(W[:, k_lo:k_lo + ncols])
    ACTS[act](xt).backward(up)
    assert nerr(dev.from_dev(dX, (M, ncols)), xt.grad.numpy()) < 2e-6


def test_dgrad_multi_segment_sum(dev):
    rng = np.random.default_rng(9)
    M, N, A = 256, 256, 8
    dz = [rng.standard_normal((M, N)).astype(np.float32) for _ in range(3)]
    Ws = [rng.standard_normal((N, 40)).astype(np.float32) for _ in range(3)]
    dX = dev.zeros((M, A))
    p = _problem(M, A, dX, A)
    for d, w in zip(dz, Ws):
        _seg(p, dev.to_dev(d), N, dev.to_dev(w) + 4 * 27, 40, N)
    _run(dev, 1, [p])
    ref = sum(torch.tensor(d) @ torch.tensor(w[:, 27:27 + A]) for d, w in zip(dz, Ws)).numpy()
    assert nerr(dev.from_dev(dX, (M, A)), ref) < 2e-6


def test_wgrad_bias_and_fused_adam(dev):
    rng = np.random.default_rng(4)
    M, N, K0, K1 = 256, 256, 256, 6
    K = K0 + K1
    ld = (K + 3) // 4 * 4
    dZ = rng.standard_normal((M, N)).astype(np.float32) / M
    x0 = rng.standard_normal((M, K0)).astype(np.float32)
    x1 = rng.standard_normal((M, K1)).astype(np.float32)
    W = np.zeros((N, ld), np.float32)
    W[:, :K] = rng.standard_normal((N, K)).astype(np.float32) * 0.05
    b = rng.standard_normal(N).astype(np.float32) * 0.05
    blob = np.concatenate([W.reshape(-1), b])
    steps = 3
    p_d, m_d, v_d, g_d = dev.to_dev(blob), dev.zeros(blob.shape), dev.zeros(blob.shape), dev.zeros(blob.shape)
    wt = torch.tensor(W[:, :K].copy(), requires_grad=True)
    bt = torch.tensor(b.copy(), requires_grad=True)
    opt = torch.optim.Adam([wt, bt], lr=3e-4)
    dZd, x0d, x1d = dev.to_dev(dZ), dev.to_dev(x0), dev.to_dev(x1)
    for t in range(1, steps + 1):
        import math
        sc = _lib.AdamScalars(w1=1 - 0.9, beta2=0.999, w2=1 - 0.999, bc2_sqrt=math.sqrt(1 - 0.999 ** t), neg_step=-(3e-4 / (1 - 0.9 ** t)), eps=1e-8, step=t)
        sc_d = dev.to_dev(np.frombuffer(bytes(sc), np.uint8))
        probs = []
        for xd, kk, off in ((x0d, K0, 0), (x1d, K1, K0)):
            q = _problem(N, kk, g_d + 4 * off, ld)
            _seg(q, dZd, N, xd, kk, M)
            q.P, q.M, q.V, q.p_ld, q.adam = p_d + 4 * off, m_d + 4 * off, v_d + 4 * off, ld, sc_d
            probs.append(q)
        qb = _problem(N, 1, g_d + 4 * N * ld, 1)
        _seg(qb, dZd, N, None, 0, M, ones=1)
        qb.P, qb.M, qb.V, qb.p_ld, qb.adam = p_d + 4 * N * ld, m_d + 4 * N * ld, v_d + 4 * N * ld, 1, sc_d
        probs.append(qb)
        _run(dev, 2, probs)
        opt.zero_grad()
        xin = torch.tensor(np.concatenate([x0, x1], 1))
        (F.linear(xin, wt, bt) * torch.tensor(dZ)).sum().backward()
        g = dev.from_dev(g_d, blob.shape)
        assert nerr(g[: N * ld].reshape(N, ld)[:, :K], wt.grad.numpy()) < 2e-6
        assert nerr(g[N * ld:], bt.grad.numpy()) < 2e-6
        opt.step()
    out = dev.from_dev(p_d, blob.shape)
    # after 3 Adam steps with (1e-7-different) gradients: normwise bound, see SURVEY.md 7.3-1
    assert nerr(out[: N * ld].reshape(N, ld)[:, :K], wt.detach().numpy()) < 2e-5
    assert np.all(out[: N * ld].reshape(N, ld)[:, K:] == 0)


def test_stacked_groups(dev):
    rng = np.random.default_rng(6)
    G, M, N, K = 3, 64, 48, 100
    x = rng.standard_normal((G, M, K)).astype(np.float32)
    W = rng.standard_normal((G, N, K)).astype(np.float32)
    b = rng.standard_normal((G, N)).astype(np.float32)
    y = dev.zeros((G, M, N))
    p = _problem(M, N, y, N)
    _seg(p, dev.to_dev(x), K, dev.to_dev(W), K, K)
    p.seg[0].a_gs, p.seg[0].b_gs, p.c_gs, p.bias, p.bias_gs = M * K, N * K, M * N, dev.to_dev(b), N
    _run(dev, 0, [p], groups=G)
    ref = torch.einsum("gmk,gnk->gmn", torch.tensor(x), torch.tensor(W)).numpy() + b[:, None, :]
    assert nerr(dev.from_dev(y, (G, M, N)), ref) < 2e-6


def test_avgl1norm_forward_backward(dev):
    rng = np.random.default_rng(2)
    rows, n = 300, 256
    x = rng.standard_normal((rows, n)).astype(np.float32)
    x[5] = 0.0  # clamp branch: mean|x| < eps
    dy = rng.standard_normal((rows, n)).astype(np.float32)
    xd, dyd, y, dx = dev.to_dev(x), dev.to_dev(dy), dev.zeros((rows, n)), dev.zeros((rows, n))
    dev.check(dev.lib.b200rl_op_avgl1norm(0, xd, None, y, rows, n, None))
    dev.check(dev.lib.b200rl_op_avgl1norm(1, xd, dyd, dx, rows, n, None))
    xt = torch.tensor(x, requires_grad=True)
    yt = xt / xt.abs().mean(-1, keepdim=True).clamp(min=1e-8)
    yt.backward(torch.tensor(dy))
    assert nerr(dev.from_dev(y, (rows, n)), yt.detach().numpy()) < 2e-6
    assert nerr(dev.from_dev(dx, (rows, n))[:5], xt.grad.numpy()[:5]) < 2e-6
    assert nerr(dev.from_dev(dx, (rows, n))[6:], xt.grad.numpy()[6:]) < 2e-6


def test_adam_matches_torch_given_identical_gradients(dev):
    rng = np.random.default_rng(8)
    n = 100_003
    p = rng.standard_normal(n).astype(np.float32)
    pd, md, vd = dev.to_dev(p), dev.zeros(n), dev.zeros(n)
    pt = torch.tensor(p.copy(), requires_grad=True)
    opt = torch.optim.Adam([pt], lr=3e-4)
    for t in range(1, 6):
        g = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0, n)).astype(np.float32)
        dev.check(dev.lib.b200rl_op_adam(pd, md, vd, dev.to_dev(g), n, 3e-4, 0.9, 0.999, 1e-8, t, None))
        pt.grad = torch.tensor(g)
        opt.step()
    st = opt.state[pt]
    assert np.array_equal(dev.from_dev(md, n), st["exp_avg"].numpy())
    assert np.array_equal(dev.from_dev(vd, n), st["exp_avg_sq"].numpy())
    # p: torch's vectorised sqrt (MKL VML) is not correctly rounded on ~0.6% of elements -> 1 ulp of the step
    assert np.abs(dev.from_dev(pd, n) - pt.detach().numpy()).max() <= 1e-6 * 3e-4 * 50


def test_polyak_bit_exact(dev):
    rng = np.random.default_rng(1)
    n = 70_001
    q, t = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    td = dev.to_dev(t)
    dev.check(dev.lib.b200rl_op_polyak(td, dev.to_dev(q), n, 0.005, None))
    dev.check(dev.lib.b200rl_stream_sync(None))
    ref = 0.005 * torch.tensor(q) + torch.tensor(t) * (1 - 0.005)
    assert np.array_equal(dev.from_dev(td, n), ref.numpy())
    # aliased source == target: TD3's Polyak of the policy onto itself
    pd = dev.to_dev(q)
    dev.check(dev.lib.b200rl_op_polyak(pd, pd, n, 0.005, None))
    dev.check(dev.lib.b200rl_stream_sync(None))
    qt = torch.tensor(q)
    assert np.array_equal(dev.from_dev(pd, n), (0.005 * qt + qt * (1 - 0.005)).numpy())


def test_sac_sample_forward_backward(dev):
    from oracle.agents import gaussian_tanh_sample

    rng = np.random.default_rng(3)
    B, A = 256, 6
    out = (rng.standard_normal((B, 2 * A)) * 0.7).astype(np.float32)
    out[:, A:] = rng.uniform(-3.0, 0.3, (B, A)).astype(np.float32)  # log_std range seen in training
    out[3, A + 1], out[4, A + 2] = 5.0, -30.0  # clamped log_std entries
    eps = rng.standard_normal((B, A)).astype(np.float32)
    da = rng.standard_normal((B, A)).astype(np.float32)
    dlp = rng.standard_normal(B).astype(np.float32)
    od, ed = dev.to_dev(out), dev.to_dev(eps)
    ad, lpd, dod = dev.zeros((B, A)), dev.zeros(B), dev.zeros((B, 2 * A))
    dev.check(dev.lib.b200rl_op_sac_sample(od, ed, ad, lpd, B, A, -20.0, 2.0, None))
    dev.check(dev.lib.b200rl_op_sac_sample_bwd(od, ed, ad, dev.to_dev(da), dev.to_dev(dlp), dod, B, A, -20.0, 2.0, None))
    dev.check(dev.lib.b200rl_stream_sync(None))
    ot = torch.tensor(out, requires_grad=True)
    mean, ls = ot.chunk(2, -1)
    a, lp = gaussian_tanh_sample(mean, torch.clamp(ls, -20.0, 2.0), torch.tensor(eps))
    ((a * torch.tensor(da)).sum() + (lp.view(-1) * torch.tensor(dlp)).sum()).backward()
    assert np.abs(dev.from_dev(ad, (B, A)) - a.detach().numpy()).max() < 2e-6
    assert nerr(dev.from_dev(lpd, B), lp.detach().numpy().reshape(-1)) < 5e-6
    # gradient: the kernel uses the analytically cancelled form (d logN/d mean = 0, d logN/d std = -1/std)
    # while fp32 autograd carries the cancellation error, so both are judged against float64 autograd
    o64 = torch.tensor(out, dtype=torch.float64, requires_grad=True)
    m64, l64 = o64.chunk(2, -1)
    a64, lp64 = gaussian_tanh_sample(m64, torch.clamp(l64, -20.0, 2.0), torch.tensor(eps, dtype=torch.float64))
    ((a64 * torch.tensor(da, dtype=torch.float64)).sum() + (lp64.view(-1) * torch.tensor(dlp, dtype=torch.float64)).sum()).backward()
    err_kernel = nerr(dev.from_dev(dod, (B, 2 * A)), o64.grad.numpy())
    err_torch32 = nerr(ot.grad.numpy(), o64.grad.numpy())
    assert err_kernel < 1e-5 or err_kernel <= 2 * err_torch32, (err_kernel, err_torch32)
    assert nerr(dev.from_dev(dod, (B, 2 * A)), ot.grad.numpy()) < 1e-4
