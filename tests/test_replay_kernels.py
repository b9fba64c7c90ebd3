"""Replay kernels through the C ABI vs the oracle and the reference-recorded golden vectors.

Bit-exact: sampled indices, gathered rows, priority scatter, running max, ring cursor.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import replay as oreplay

S, A = 11, 3  # Hopper-v4


def _fill(dev, h, n, seed=3, group=0):
    rows = oreplay.synthetic_rows(n, S, A, seed)
    packed = dev.pack_rows(h, rows["state"], rows["action"], rows["reward"], rows["next_state"], rows["done"])
    dev.replay_append(h, packed, group)
    return rows


def test_lap_sampling_matches_reference_golden(dev):
    g = gu.load("replay_cases")
    cap = 3000
    h = dev.replay_create(cap, S, A)
    rows = _fill(dev, h, cap)
    for nm in ("ones", "rand", "zeros_mixed", "pow"):
        for size in (cap, 1777, 1):
            key = f"samp/{nm}/{size}"
            pri, u, ind = g[key + "/priority"], g[key + "/u"], g[key + "/ind"]
            dev.replay_write_priority(h, pri)
            dev.check(dev.lib.b200rl_replay_set_cursor(h, 0, size % cap, size))
            idx, batch, _ = dev.replay_sample(h, 256, S, A, u=u)
            assert np.array_equal(idx[0], ind), key
            assert np.array_equal(batch["state"][0], rows["state"][ind]), key
            assert np.array_equal(batch["next_state"][0], rows["next_state"][ind]), key
            assert np.array_equal(batch["action"][0], rows["action"][ind]), key
            assert np.array_equal(batch["reward"][0], rows["reward"][ind]), key
            assert np.array_equal(batch["done"][0], rows["done"][ind]), key
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_uniform_sampling_matches_reference_golden(dev):
    g = gu.load("replay_cases")
    cap = 3000
    h = dev.replay_create(cap, S, A, prioritized=False)
    _fill(dev, h, cap)
    for size in (cap, 1234, 1):
        dev.check(dev.lib.b200rl_replay_set_cursor(h, 0, size % cap, size))
        idx, _, _ = dev.replay_sample(h, 256, S, A, u=g[f"uni/{size}/u"])
        assert np.array_equal(idx[0], g[f"uni/{size}/ind"])
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_lap_sampling_one_million_rows(dev):
    """The size BASELINE.json quotes: 1M priorities, indices bit-exact vs the reference."""
    g = gu.load("replay_cases")
    big = (np.maximum(np.abs(np.random.default_rng(11).standard_normal(1_000_000) * 3), 1.0) ** 0.4).astype(np.float32)
    h = dev.replay_create(1_000_000, 2, 1)
    dev.replay_write_priority(h, big)
    dev.check(dev.lib.b200rl_replay_set_cursor(h, 0, 0, 1_000_000))
    idx, _, _ = dev.replay_sample(h, 256, 2, 1, u=g["big/u"])
    assert np.array_equal(idx[0], g["big/ind"])
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_append_ring_wrap_scatter_and_running_max(dev):
    g = gu.load("replay_cases")
    cap = 3000
    h = dev.replay_create(cap, S, A)
    ref = oreplay.OracleLAP(cap, "Hopper-v4")
    n = 3500
    act_norm = (g["wrap/act"] / ref.action_scale - ref.action_bias).astype(np.float32)
    packed = dev.pack_rows(h, g["wrap/obs"], act_norm, g["wrap/rew"], g["wrap/next_obs"], g["wrap/notdone"])
    # one row at a time up to the mid-stream priority update, then in bulk across the wrap point
    for i in range(1701):
        dev.replay_append(h, packed[i:i + 1])
        ref.append([g["wrap/obs"][i], g["wrap/act"][i], float(g["wrap/rew"][i]), g["wrap/next_obs"][i], float(g["wrap/notdone"][i])])
    idx, _, _ = dev.replay_sample(h, 64, S, A, u=g["wrap/mid_u"])
    assert np.array_equal(idx[0], g["wrap/mid_ind"])
    ref.sample(64, u=torch.tensor(g["wrap/mid_u"]))
    pr = dev.to_dev(g["wrap/mid_prio"])
    dev.check(dev.lib.b200rl_replay_update_priority(h, pr, 64, None))
    ref.update_priority(torch.tensor(g["wrap/mid_prio"]))
    dev.replay_append(h, packed[1701:n])
    for i in range(1701, n):
        ref.append([g["wrap/obs"][i], g["wrap/act"][i], float(g["wrap/rew"][i]), g["wrap/next_obs"][i], float(g["wrap/notdone"][i])])
    ptr, size, maxp = dev.replay_state(h)
    assert [ptr, size, ptr] == list(g["wrap/ptr_size_len"])
    assert maxp == float(g["wrap/max_priority"])
    assert np.array_equal(dev.replay_priority(h, cap), g["wrap/priority"])
    stride = dev.lib.b200rl_replay_row_stride(h)
    rows = np.empty((cap, stride), np.float32)
    dev.check(dev.lib.b200rl_replay_read_rows(h, 0, 0, cap, rows.ctypes.data, None))
    o = dev.row_offsets(h)
    assert np.array_equal(rows[:, o["state"]:o["state"] + S], ref.state.astype(np.float32))
    assert np.array_equal(rows[:, o["action"]:o["action"] + A], ref.action.astype(np.float32))
    assert np.array_equal(rows[:, o["next_state"]:o["next_state"] + S], ref.next_state.astype(np.float32))
    assert o["next_state"] % 4 == 0 and o["state"] == 0  # both observation vectors on 16-byte boundaries
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_incrementally_repaired_sum_tree_tracks_the_oracle(dev):
    """The sampler's sum tree is repaired in place by append / update_priority; after many rounds of
    {append a few rows (crossing tile, super-tile and ring boundaries), sample, write priorities back}
    the indices must still be the reference's (torch.cumsum + searchsorted on the whole array)."""
    cap = 40_000  # 157 tiles, 3 super-tiles; not a multiple of the tile size
    h = dev.replay_create(cap, S, A)
    ref = oreplay.OracleLAP(cap, "Hopper-v4")
    rng = np.random.default_rng(77)

    def rows(n):
        obs, act = rng.standard_normal((n, S)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32)
        rew, nobs, nd = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, S)).astype(np.float32), (rng.random(n) > 0.1).astype(np.float32)
        for i in range(n):
            ref.append([obs[i], act[i], float(rew[i]), nobs[i], float(nd[i])])
        return dev.pack_rows(h, obs, (act / ref.action_scale - ref.action_bias).astype(np.float32), rew, nobs, nd)

    dev.replay_append(h, rows(39_000))  # bulk: invalidates the tree, the first sample rebuilds it
    B = 96
    for rnd in range(40):
        n_new = [1, 3, 300, 1, 700, 5][rnd % 6]  # from round ~2 on the ring wraps
        dev.replay_append(h, rows(n_new))
        u = rng.random(B).astype(np.float32)
        if rnd % 7 == 0:
            u[:3] = [0.0, np.nextafter(np.float32(1), np.float32(0)), 0.5]
        idx, _, _ = dev.replay_sample(h, B, S, A, u=u)
        ref.sample(B, u=torch.tensor(u))
        assert np.array_equal(idx[0], ref.ind), f"round {rnd}"
        pr = (np.maximum(np.abs(rng.standard_normal(B)) * 4, 1.0) ** 0.4).astype(np.float32)
        dev.check(dev.lib.b200rl_replay_update_priority(h, dev.to_dev(pr), B, None))
        ref.update_priority(torch.tensor(pr))
    assert np.array_equal(dev.replay_priority(h, cap), ref.priority.numpy().astype(np.float32))
    assert dev.replay_state(h)[2] == float(ref.max_priority)
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_duplicate_scatter_last_wins_and_reset(dev):
    g = gu.load("replay_cases")
    cap = 3000
    h = dev.replay_create(cap, S, A)
    _fill(dev, h, cap)
    dev.replay_write_priority(h, np.ones(cap, np.float32), max_priority=1.0)
    # force the "last sampled indices" through a sample whose draws hit exactly those rows:
    # with all-ones priorities, u = (i + 0.5) / cap selects row i
    want = g["dup/ind"]
    u = ((want + 0.5) / cap).astype(np.float32)
    idx, _, _ = dev.replay_sample(h, len(want), S, A, u=u)
    assert np.array_equal(idx[0], want)
    pr = dev.to_dev(g["dup/prio"])
    dev.check(dev.lib.b200rl_replay_update_priority(h, pr, len(want), None))
    assert np.array_equal(dev.replay_priority(h, cap), g["dup/priority_after"])
    assert dev.replay_state(h)[2] == float(g["dup/max_priority"])
    pri = dev.replay_priority(h, cap)
    pri[100] = 1.0
    dev.replay_write_priority(h, pri)
    dev.check(dev.lib.b200rl_replay_reset_max_priority(h, None))
    assert dev.replay_state(h)[2] == float(g["dup/max_after_reset"])
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_stacked_seeds_share_rows_and_device_rng(dev):
    """groups > 1 with shared transitions: each seed samples with its own priorities; the
    on-device Philox draws stay in range and differ between seeds."""
    cap, G = 5000, 3
    h = dev.replay_create(cap, S, A, groups=G, share_rows=True)
    rows = None
    for gidx in range(G):
        rows = _fill(dev, h, cap, group=gidx)
    rng = np.random.default_rng(5)
    pris = [rng.uniform(1, 8, cap).astype(np.float32) for _ in range(G)]
    for gidx in range(G):
        dev.replay_write_priority(h, pris[gidx], group=gidx)
    u = rng.random((G, 128), dtype=np.float32)
    idx, batch, _ = dev.replay_sample(h, 128, S, A, u=u, groups=G)
    for gidx in range(G):
        assert np.array_equal(idx[gidx], oreplay.sample_indices_numpy(pris[gidx], u[gidx]))
        assert np.array_equal(batch["state"][gidx], rows["state"][idx[gidx]])
    idx2, _, _ = dev.replay_sample(h, 128, S, A, u=None, groups=G)
    assert idx2.min() >= 0 and idx2.max() < cap
    assert not np.array_equal(idx2[0], idx2[1])
    dev.check(dev.lib.b200rl_replay_destroy(h))


@pytest.mark.parametrize("dims", [(376, 17), (64, 6), (100, 8), (27, 8), (17, 6), (130, 2)])
def test_row_gather_every_width(dev, dims):
    """The gather has a 128-bit path (16-byte aligned batch rows of >= 64 floats, with aligned and unaligned sources)
    and a lane-per-float path: every env shape of BASELINE.json plus widths that land on each branch, all four
    destination alignments (batch rows of odd width start at every offset mod 16 bytes), bit-exact, stacked seeds."""
    s, a = dims
    cap, G, B = 700, 2, 67
    h = dev.replay_create(cap, s, a, groups=G, share_rows=True)
    rows = oreplay.synthetic_rows(cap, s, a, seed=s)
    packed = dev.pack_rows(h, rows["state"], rows["action"], rows["reward"], rows["next_state"], rows["done"])
    for g in range(G):
        dev.replay_append(h, packed, g)
    pri = np.random.default_rng(s).uniform(1.0, 9.0, (G, cap)).astype(np.float32)
    for g in range(G):
        dev.replay_write_priority(h, pri[g], group=g)
    u = np.random.default_rng(a).random((G, B), dtype=np.float32)
    idx, batch, _ = dev.replay_sample(h, B, s, a, u=u, groups=G)
    for g in range(G):
        assert np.array_equal(idx[g], oreplay.sample_indices_numpy(pri[g], u[g]))
        for k in ("state", "action", "reward", "next_state", "done"):
            assert np.array_equal(batch[k][g], rows[k][idx[g]]), (dims, g, k)
    dev.check(dev.lib.b200rl_replay_destroy(h))


def test_lap_sampling_beyond_two_million_rows(dev):
    """More than 128 super-tiles (capacity > 2 M rows): a lane owns five or more super sums and takes the
    re-reading branch of the top tree level instead of the register-resident one; indices bit-exact vs the oracle."""
    n = 2_200_000
    pri = (np.maximum(np.abs(np.random.default_rng(21).standard_normal(n) * 3), 1.0) ** 0.4).astype(np.float32)
    u = np.random.default_rng(22).random(256, dtype=np.float32)
    u[:3] = (0.0, 1.0 - 2.0 ** -24, 0.5)
    h = dev.replay_create(n, 2, 1)
    dev.replay_write_priority(h, pri)
    dev.check(dev.lib.b200rl_replay_set_cursor(h, 0, 0, n))
    idx, _, _ = dev.replay_sample(h, 256, 2, 1, u=u)
    assert np.array_equal(idx[0], oreplay.sample_indices_numpy(pri, u))
    dev.check(dev.lib.b200rl_replay_destroy(h))
