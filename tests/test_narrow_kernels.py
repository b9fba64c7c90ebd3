"""narrow.cuh -- the problems of a stacked-seed update that are not 256-wide contractions: weight gradients a few columns wide
(with their layer's bias gradient and Adam), few-output heads as transposed problems, outputs up to 32 columns wide, and the
rank-1 dgrad through a one-output layer.  Plain fp32 kernels (no TF32): held to the fp32 op bar (2e-6 normwise) against torch.
They run on the GPU (`-m gpu`, with the tensor-core dispatch forced on) and, on the GPU-less container, under the host
emulation with its guard-page allocator (B200RL_NARROW=1 routes any seed count through them)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from sac_td3_td7_b200 import _lib
from test_op_kernels import _problem, _run, _seg, nerr


@pytest.fixture()
def nar(dev, monkeypatch):
    monkeypatch.setenv("B200RL_NARROW", "1")
    if dev.backend == "cuda":
        monkeypatch.setenv("B200RL_TC", "1")
    return dev

FP32_TOL = 2e-6


def _adam_scalars(dev, t=1, lr=3e-4):
    sc = _lib.AdamScalars(w1=1 - 0.9, beta2=0.999, w2=1 - 0.999, bc2_sqrt=math.sqrt(1 - 0.999 ** t), neg_step=-(lr / (1 - 0.9 ** t)), eps=1e-8, step=t)
    return dev.to_dev(np.frombuffer(bytes(sc), np.uint8))


@pytest.mark.parametrize("G,K0,K1", [(1, 27, 8), (3, 17, 6), (2, 40, 17)])
def test_narrow_wgrad_with_folded_bias_and_adam(nar, G, K0, K1):
    """First-layer weight gradient of a critic: [state | action] column blocks (the action block starts at the next
    multiple of 4), each a narrow problem; the layer's bias gradient rides with the first; Adam in place."""
    dev = nar
    rng = np.random.default_rng(G * 100 + K0)
    M, N = 256, 256
    off1 = (K0 + 3) // 4 * 4
    ld = off1 + (K1 + 3) // 4 * 4
    blob_n = N * ld + N
    dZ = rng.standard_normal((G, M, N)).astype(np.float32) / M
    x0 = rng.standard_normal((G, M, K0)).astype(np.float32)
    x1 = rng.standard_normal((G, M, K1)).astype(np.float32)
    blob = np.zeros((G, blob_n), np.float32)
    W = blob[:, : N * ld].reshape(G, N, ld)
    W[:, :, :K0] = rng.standard_normal((G, N, K0)) * 0.05
    W[:, :, off1:off1 + K1] = rng.standard_normal((G, N, K1)) * 0.05
    blob[:, N * ld:] = rng.standard_normal((G, N)) * 0.05
    p_d, m_d, v_d, g_d = dev.to_dev(blob), dev.zeros(blob.shape), dev.zeros(blob.shape), dev.zeros(blob.shape)
    dZd, x0d, x1d, sc_d = dev.to_dev(dZ), dev.to_dev(x0), dev.to_dev(x1), _adam_scalars(dev)
    probs = []
    for xd, kk, off in ((x0d, K0, 0), (x1d, K1, off1)):
        q = _problem(N, kk, g_d + 4 * off, ld)
        _seg(q, dZd, N, xd, kk, M)
        q.seg[0].a_gs, q.seg[0].b_gs, q.c_gs, q.p_gs = M * N, M * kk, blob_n, blob_n
        q.P, q.M, q.V, q.p_ld, q.adam = p_d + 4 * off, m_d + 4 * off, v_d + 4 * off, ld, sc_d
        probs.append(q)
    qb = _problem(N, 1, g_d + 4 * N * ld, 1)
    _seg(qb, dZd, N, None, 0, M, ones=1)
    qb.seg[0].a_gs, qb.c_gs, qb.p_gs = M * N, blob_n, blob_n
    qb.P, qb.M, qb.V, qb.p_ld, qb.adam = p_d + 4 * N * ld, m_d + 4 * N * ld, v_d + 4 * N * ld, 1, sc_d
    probs.append(qb)
    _run(dev, 2, probs, groups=G)
    g = dev.from_dev(g_d, blob.shape)
    out, mo = dev.from_dev(p_d, blob.shape), dev.from_dev(m_d, blob.shape)
    for s in range(G):
        gw = g[s, : N * ld].reshape(N, ld)
        ref0 = (torch.tensor(dZ[s]).T @ torch.tensor(x0[s])).numpy()
        ref1 = (torch.tensor(dZ[s]).T @ torch.tensor(x1[s])).numpy()
        assert nerr(gw[:, :K0], ref0) < FP32_TOL and nerr(gw[:, off1:off1 + K1], ref1) < FP32_TOL
        assert nerr(g[s, N * ld:], dZ[s].sum(0)) < FP32_TOL
        assert np.all(gw[:, K0:off1] == 0) and np.all(out[s, : N * ld].reshape(N, ld)[:, K0:off1] == 0)  # padding untouched
        # Adam applied to the kernel's own gradient is exact arithmetic on it
        wt = torch.tensor(blob[s].copy(), requires_grad=True)
        opt = torch.optim.Adam([wt], lr=3e-4)
        wt.grad = torch.tensor(g[s])
        opt.step()
        live = np.ones(blob_n, bool)
        live[: N * ld].reshape(N, ld)[:, K0:off1] = False
        live[: N * ld].reshape(N, ld)[:, off1 + K1:] = False
        assert np.abs(out[s] - wt.detach().numpy())[live].max() <= 2e-7
        assert np.array_equal(mo[s][live], opt.state[wt]["exp_avg"].numpy()[live])


def test_narrow_wgrad_transposed_few_outputs(nar):
    """Weight gradient of a head with a few outputs (action head: 8 x 256): computed as its transpose."""
    dev = nar
    rng = np.random.default_rng(11)
    G, M, N, K = 2, 256, 8, 256
    dZ = rng.standard_normal((G, M, N)).astype(np.float32) / M
    x = rng.standard_normal((G, M, K)).astype(np.float32)
    g_d = dev.zeros((G, N, K))
    q = _problem(N, K, g_d, K)
    _seg(q, dev.to_dev(dZ), N, dev.to_dev(x), K, M)
    q.seg[0].a_gs, q.seg[0].b_gs, q.c_gs = M * N, M * K, N * K
    _run(dev, 2, [q], groups=G)
    g = dev.from_dev(g_d, (G, N, K))
    for s in range(G):
        assert nerr(g[s], (torch.tensor(dZ[s]).T @ torch.tensor(x[s])).numpy()) < FP32_TOL


@pytest.mark.parametrize("A", [6, 8, 17])
def test_narrow_rows_forward_heads_and_action_gradient(nar, A):
    """Action head with target-policy smoothing noise, scalar head, and the gradient with respect to the action (three
    dgrad terms, tanh derivative): outputs <= 32 columns wide, eight lanes per row, passes of eight columns."""
    dev = nar
    rng = np.random.default_rng(A)
    G, M, K = 3, 256, 256
    x = rng.standard_normal((G, M, K)).astype(np.float32)
    Wa = (rng.standard_normal((G, A, K)) / 16).astype(np.float32)
    ba = rng.standard_normal((G, A)).astype(np.float32)
    noise = rng.standard_normal((G, M, A)).astype(np.float32)
    wq = (rng.standard_normal((G, 1, K)) / 16).astype(np.float32)
    ya, yq = dev.zeros((G, M, A)), dev.zeros((G, M, 1))
    xd = dev.to_dev(x)
    pa = _problem(M, A, ya, A)
    _seg(pa, xd, K, dev.to_dev(Wa), K, K)
    pa.seg[0].a_gs, pa.seg[0].b_gs, pa.c_gs = M * K, A * K, M * A
    pa.bias, pa.bias_gs, pa.act = dev.to_dev(ba), A, 4
    pa.aux, pa.aux_ld, pa.aux_gs, pa.noise_sigma, pa.noise_clip = dev.to_dev(noise), A, M * A, 0.2, 0.5
    pq = _problem(M, 1, yq, 1)
    _seg(pq, xd, K, dev.to_dev(wq), K, K)
    pq.seg[0].a_gs, pq.seg[0].b_gs, pq.c_gs = M * K, K, M
    _run(dev, 0, [pa, pq], groups=G)
    got_a, got_q = dev.from_dev(ya, (G, M, A)), dev.from_dev(yq, (G, M, 1))
    for s in range(G):
        xt = torch.tensor(x[s])
        ref = (torch.tanh(F.linear(xt, torch.tensor(Wa[s]), torch.tensor(ba[s]))) + (torch.tensor(noise[s]) * 0.2).clamp(-0.5, 0.5)).clamp(-1, 1).numpy()
        assert nerr(got_a[s], ref) < FP32_TOL
        assert nerr(got_q[s], F.linear(xt, torch.tensor(wq[s])).numpy()) < FP32_TOL
    # d a = (dz1 W1[:, k1:k1+A] + dz2 W2[:, k2:k2+A] + dz3 W3[:, k3:k3+A]) * tanh'(a)
    lds, klo = (48, 48, 280), (28, 28, 256)
    dzs = [rng.standard_normal((G, M, K)).astype(np.float32) for _ in range(3)]
    Ws = [(rng.standard_normal((G, K, ld)) / 16).astype(np.float32) for ld in lds]
    h = np.tanh(rng.standard_normal((G, M, A))).astype(np.float32)
    da = dev.zeros((G, M, A))
    pd = _problem(M, A, da, A)
    for dz, w, ld, k in zip(dzs, Ws, lds, klo):
        _seg(pd, dev.to_dev(dz), K, dev.to_dev(w) + 4 * k, ld, K)
        sg = pd.seg[pd.nseg - 1]
        sg.a_gs, sg.b_gs = M * K, K * ld
    pd.c_gs, pd.aux, pd.aux_ld, pd.aux_gs, pd.act = M * A, dev.to_dev(h), A, M * A, 3
    _run(dev, 1, [pd], groups=G)
    got = dev.from_dev(da, (G, M, A))
    for s in range(G):
        ref = sum(torch.tensor(dz[s]) @ torch.tensor(w[s][:, k:k + A]) for dz, w, k in zip(dzs, Ws, klo)).numpy() * (1 - h[s] ** 2)
        assert nerr(got[s], ref) < FP32_TOL


def test_rank1_dgrad_through_a_one_output_layer(nar):
    dev = nar
    rng = np.random.default_rng(13)
    G, M, N = 3, 256, 256
    dq = rng.standard_normal((G, M, 1)).astype(np.float32)
    w = rng.standard_normal((G, 1, N)).astype(np.float32)
    h = F.elu(torch.tensor(rng.standard_normal((G, M, N)).astype(np.float32))).numpy()
    dX = dev.zeros((G, M, N))
    p = _problem(M, N, dX, N)
    _seg(p, dev.to_dev(dq), 1, dev.to_dev(w), N, 1)
    p.seg[0].a_gs, p.seg[0].b_gs, p.c_gs = M, N, M * N
    p.aux, p.aux_ld, p.aux_gs, p.act = dev.to_dev(h), N, M * N, 2
    _run(dev, 1, [p], groups=G)
    got = dev.from_dev(dX, (G, M, N))
    for s in range(G):
        ref = (dq[s] * w[s]) * np.where(h[s] > 0, 1.0, h[s] + 1.0)
        assert nerr(got[s], ref) < FP32_TOL
