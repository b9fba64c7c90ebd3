"""Statistical tests of the device RNG the production path (train_ops_n) runs on: Philox4x32-10 uniforms of the replay
sampler (the reference draws torch.rand, rl/replay_memory/lap.py:50) and Box-Muller Gaussians of the target-smoothing /
reparameterisation noise (torch.randn_like, rl/agent/td7.py:190-193, sac.py:158-171).  Round 1 only range-checked them.

The uniforms are observed through the sampler itself: on a UNIFORM buffer of exactly 2^24 rows the sampled index is
ceil(u * 2^24) - 1, i.e. the 24-bit integer the generator produced (0 and 1 share index 0)."""
import ctypes as C

import numpy as np
import pytest
import torch
from scipy import stats

pytestmark = pytest.mark.gpu

N24 = 1 << 24


def _idx(lib, h, G, B):
    out = np.empty((G, B), np.int64)
    assert lib.b200rl_replay_last_indices(h, out.ctypes.data, None) == 0
    return out


def test_sampler_uniforms_are_uniform_independent_and_below_one():
    from sac_td3_td7_b200 import _lib

    lib = _lib.lib()
    G, B, calls = 2, 4096, 128  # 2^20 draws per stacked seed
    h = C.c_void_p()
    _lib.check(lib.b200rl_replay_create(N24, 1, 1, G, 0, 1, C.byref(h)))  # uniform buffer, one shared (all-zero) row store
    for g in range(G):
        _lib.check(lib.b200rl_replay_set_cursor(h, g, 0, N24))
    outs = [torch.empty(G * B * d, device="cuda") for d in (1, 1, 1, 1, 1)]
    draws = []
    for _ in range(calls):
        _lib.check(lib.b200rl_replay_sample(h, B, None, *[o.data_ptr() for o in outs], None))
        draws.append(_idx(lib, h, G, B))
    k = np.stack(draws, 1).reshape(G, -1)  # [G, calls * B]
    assert k.min() >= 0 and k.max() <= N24 - 1, "a uniform reached 1.0 (index past the last row)"
    u = (k + 0.5) / N24
    for g in range(G):
        ks = stats.kstest(u[g], "uniform")
        assert ks.pvalue > 1e-3, f"seed {g}: KS vs U(0,1) p = {ks.pvalue:.2e} (D = {ks.statistic:.2e})"
        n = u[g].size
        assert abs(u[g].mean() - 0.5) < 5 * np.sqrt(1 / 12 / n)
        assert abs(u[g].var() - 1 / 12) < 5 * np.sqrt(1 / 180 / n)
        # serial independence: lag-1 correlation and the correlation between consecutive calls
        x = u[g] - 0.5
        assert abs(np.mean(x[1:] * x[:-1]) * 12) < 5 / np.sqrt(n)
        # low bits: the 24-bit integers are uniform mod 256 (chi-square over 256 cells)
        chi = stats.chisquare(np.bincount((k[g] & 255).astype(np.int64), minlength=256))
        assert chi.pvalue > 1e-4, chi
        # no call repeats an earlier call's draws (counter advances)
        per_call = k[g].reshape(calls, B)
        assert len({per_call[i].tobytes() for i in range(calls)}) == calls
    # per-seed streams are distinct and uncorrelated
    assert (k[0] == k[1]).mean() < 1e-3
    assert abs(np.corrcoef(u[0], u[1])[0, 1]) < 5 / np.sqrt(u[0].size)
    lib.b200rl_replay_destroy(h)


def test_gaussian_noise_moments_streams_and_continuity():
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env, G, B, A = "Humanoid-v4", 2, 1024, 17
    ag = TD7(env, use_lap=True, n_seeds=G, batch_size=B)
    ag.seed_population(3)
    buf = LAPReplayMemory(4096, env, n_seeds=G, share_rows=True)
    buf.fill_synthetic(4096, seed=0)
    z = np.empty((G, 62, B * A), np.float32)
    inds = []
    tmp = np.empty(B * A, np.float32)
    for t in range(62):
        if t == 31:  # re-capture the update graphs mid-run (re-binding a replay drops them): the stream must simply continue
            ag._lib.b200rl_agent_set_rng(ag._h, 0xB200B200)
        ag.train_ops_n(1, buf)
        inds.append(buf.ind.copy())
        for g in range(G):
            assert ag._lib.b200rl_agent_debug_read(ag._h, 5, g, tmp.ctypes.data, None) == 0
            z[g, t] = tmp
    for g in range(G):
        x = z[g].reshape(-1).astype(np.float64)
        n = x.size  # 1 079 296 draws
        assert np.isfinite(x).all()
        assert abs(x.mean()) < 5 / np.sqrt(n)
        assert abs(x.var() - 1.0) < 5 * np.sqrt(2.0 / n)
        assert abs(stats.skew(x)) < 5 * np.sqrt(6.0 / n)
        assert abs(stats.kurtosis(x)) < 5 * np.sqrt(24.0 / n)
        ks = stats.kstest(x, "norm")
        assert ks.pvalue > 1e-3, f"seed {g}: KS vs N(0,1) p = {ks.pvalue:.2e}"
        assert np.abs(x).max() < 6.5  # Box-Muller with 24-bit uniforms: |z| <= sqrt(2 ln 2^24) = 5.77
        # consecutive updates never reuse a counter (also across the graph re-capture at t = 31)
        assert len({z[g, t].tobytes() for t in range(62)}) == 62
        lag = np.mean(z[g, 1:].astype(np.float64) * z[g, :-1])
        assert abs(lag) < 5 / np.sqrt(61 * B * A)
    assert abs(np.corrcoef(z[0].reshape(-1), z[1].reshape(-1))[0, 1]) < 5 / np.sqrt(z[0].size)
    # the sampler's stream inside bursts advances with the update counter: no two updates drew the same indices
    assert len({i.tobytes() for i in inds}) == 62
    assert not np.array_equal(inds[0][0], inds[0][1])  # stacked seeds sample differently
