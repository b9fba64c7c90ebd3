"""Stacked-seed populations through the drop-in API (GPU): acting / evaluation, per-step training, persistence, and the
round-2 fixes to the acting buffers, the loss mailbox after unpickling, and bursts longer than the device info ring.

Reference interfaces: ``agent.sample`` (rl/agent/td7.py:141-162, td3.py:114-141, sac.py:131-159), ``Agent.save/load``
(rl/agent/abc.py:38-55), ``run_train_ops`` + ``log_train_infos`` (rl/runner/run.py:53-96), ``test_agent``
(rl/runner/run.py:21-45), ``append`` (rl/replay_memory/lap.py:31-43)."""
import ctypes as C
import pickle
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def _oracle_action(algo, mods, obs):
    """Deterministic policy output of one seed's networks by the CPU oracle (pre-scale action; SAC: mean | log_std)."""
    from oracle import nets as onets

    s = torch.as_tensor(obs, dtype=torch.float32).reshape(1, -1)
    if algo == "TD7":
        zs = onets.encode_state(_sd(mods["fixed_encoder"]), s)
        return onets.sale_actor(_sd(mods["policy"]), s, zs)[0].numpy()
    out = onets.mlp(_sd(mods["policy"]), s)[0]
    return (torch.tanh(out) if algo == "TD3" else out).numpy()


@pytest.mark.parametrize("algo,env,G", [("TD7", "Humanoid-v4", 1), ("TD7", "Ant-v4", 5), ("TD7", "HalfCheetah-v4", 9), ("TD3", "Hopper-v4", 3), ("SAC", "Walker2d-v4", 4),
                                        ("TD3", "Humanoid-v4", 1)])
def test_acting_one_kernel_matches_oracle_for_every_stacked_seed(algo, env, G):
    """agent.sample(obs) for a population returns one action per seed, each equal (1e-5) to the oracle forward of THAT
    seed's weights; a [S] observation is broadcast; the result buffers are sized by n_seeds (round-1 overflow)."""
    from sac_td3_td7_b200.rl import agent as A

    cls = getattr(A, algo)
    torch.manual_seed(11)
    ag = cls(env, n_seeds=G)
    if G > 1:
        ag.seed_population(500)
    S, Ad = ag.state_dim, ag.action_dim
    rng = np.random.default_rng(G)
    obs = rng.standard_normal((G, S)).astype(np.float32)
    raw = ag._policy_out(obs if G > 1 else obs[0])
    raw = raw.reshape(G, -1)
    assert raw.shape == (G, Ad * (2 if algo == "SAC" else 1))
    for g in range(G):
        mods = ag.seed_modules(g)
        ref = _oracle_action(algo, mods, obs[g])
        assert np.abs(raw[g] - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (g, np.abs(raw[g] - ref).max())
    # public API: deterministic sample in environment units, per seed
    act = np.asarray(ag.sample(obs if G > 1 else obs[0], True)).reshape(G, Ad)
    for g in range(G):
        ref = _oracle_action(algo, ag.seed_modules(g), obs[g])
        ref = np.tanh(ref[:Ad]) if algo == "SAC" else np.clip(ref, -1, 1)
        assert np.abs(act[g] - (ref * ag.action_scale + ag.action_bias)).max() <= 2e-5
    # stochastic sample keeps the shape and the bounds
    st = np.asarray(ag.sample(obs if G > 1 else obs[0], False)).reshape(G, Ad)
    assert np.all(np.abs((st - ag.action_bias) / ag.action_scale) <= 1.0 + 1e-6)
    # broadcast of one observation over the population, and the device-tensor form
    if G > 1:
        b = ag._policy_out(obs[0])
        assert b.shape == raw.shape and np.abs(b[0] - raw[0]).max() <= 1e-6
        assert np.abs(b[1] - _oracle_action(algo, ag.seed_modules(1), obs[0])).max() <= 1e-5 * max(1.0, np.abs(b[1]).max())
    dv = ag.policy_out_device(torch.as_tensor(obs).cuda())
    torch.cuda.synchronize()
    assert np.abs(dv.cpu().numpy() - raw).max() <= 1e-6
    with pytest.raises(ValueError):
        ag._policy_out(np.zeros((G + 1, S), np.float32))


def test_acting_sees_the_weights_of_the_latest_update_and_is_fast():
    """The acting kernel reads the live parameter blobs in stream order: an action taken after k updates equals the oracle
    forward of the weights read back after those k updates.  Also reports the host-to-host latency."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env = "Humanoid-v4"
    torch.manual_seed(2)
    ag, buf = TD7(env, use_lap=True, target_update_rate=3), LAPReplayMemory(8192, env)
    buf.fill_synthetic(8192, seed=0)
    obs = np.random.default_rng(0).standard_normal(376).astype(np.float32)
    a0 = ag._policy_out(obs)
    ag.train_ops_n(7, buf)  # two hard syncs: the fixed encoder the policy acts through has changed
    a1 = ag._policy_out(obs)
    assert np.abs(a1 - a0).max() > 1e-4
    ref = _oracle_action("TD7", ag.seed_modules(0), obs)
    assert np.abs(a1 - ref).max() <= 1e-5
    for _ in range(50):
        ag._policy_out(obs)
    t0 = time.perf_counter()
    for _ in range(500):
        ag._policy_out(obs)
    us = (time.perf_counter() - t0) / 500 * 1e6
    print(f"TD7 Humanoid agent.sample host-to-host: {us:.1f} us")
    assert us < 200.0


def test_population_pickle_roundtrip_keeps_every_seed():
    """save()/load() of a stacked agent (round 1 persisted seed 0 only): parameters, Adam moments, value ranges and counters
    of EVERY seed survive, and the restored population continues bit-identically."""
    from sac_td3_td7_b200.rl.agent import SAC, TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    for cls, env, kw, bcls in ((TD7, "Hopper-v4", dict(use_lap=True, target_update_rate=3), LAPReplayMemory), (SAC, "Hopper-v4", {}, SimpleReplayMemory)):
        G = 3
        torch.manual_seed(5)
        ag = cls(env, n_seeds=G, **kw)
        ag.seed_population(40)
        buf = bcls(4096, env, n_seeds=G, share_rows=True)
        buf.fill_synthetic(4096, seed=3)
        ag.train_ops_n(7, buf)
        blob = pickle.dumps(ag)
        back = pickle.loads(blob)
        assert back.n_seeds == G and back._n_calls == 7
        for nid, (attr, _) in enumerate(ag.NETS):
            kinds = (0, 1, 2) if attr in ag.TRAINED else (0,)
            for w in kinds:
                a, b = ag._read_blob(nid, w, -1)[0], back._read_blob(nid, w, -1)[0]
                assert a.shape[0] == G and np.array_equal(a, b), (attr, w)
            assert not np.array_equal(ag._read_blob(nid, 0, 0)[0], ag._read_blob(nid, 0, 1)[0])  # seeds differ
        if cls is TD7:
            assert np.array_equal(ag._range(-1), back._range(-1)) and ag._range(-1).shape == (G, 4)
        else:
            assert np.array_equal(ag._extra_state()["__tmp__"], back._extra_state()["__tmp__"])
        # both continue from the same state with the same device RNG streams: identical infos for every seed.  The restored
        # population trains on a twin buffer brought to the same priorities by an identically seeded twin run.
        ag.train_ops_n(4, buf)
        r1 = ag.read_infos(8, 4, -1)
        buf_b = bcls(4096, env, n_seeds=G, share_rows=True)
        buf_b.fill_synthetic(4096, seed=3)
        torch.manual_seed(5)
        twin = cls(env, n_seeds=G, **kw)
        twin.seed_population(40)
        twin.train_ops_n(7, buf_b)
        back.train_ops_n(4, buf_b)
        r2 = back.read_infos(8, 4, -1)
        assert r1.shape == (G, 4, len(ag.INFO_KEYS))
        assert np.array_equal(r1, r2)
        if cls is TD7:
            assert torch.equal(buf.priority, buf_b.priority)


def test_loss_mailbox_works_after_unpickling():
    """b200rl_agent_set_counters used to null the mailbox pointer: every float(info[...]) then spun for the 2 s timeout."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env = "Hopper-v4"
    ag, buf = TD7(env, use_lap=True, sync_info=False), LAPReplayMemory(4096, env)
    buf.fill_synthetic(4096, seed=0)
    ag.train_ops_n(3, buf)
    back = pickle.loads(pickle.dumps(ag))
    info = back.train_ops(buf.sample(256), replay_buffer=buf)
    v = C.c_float()
    t0 = time.perf_counter()
    rc = back._lib.b200rl_agent_wait_info(back._h, back._n_calls, 1, C.byref(v))
    dt = time.perf_counter() - t0
    assert rc == 0 and dt < 0.5, (rc, dt)
    assert float(info["train/q_fn"]) == pytest.approx(v.value, rel=0, abs=0)


def test_bursts_longer_than_the_info_ring_stay_readable():
    """run_rl_w_ckpt trains n_ops = sum of episode lengths (thousands) and evaluates the returned infos afterwards
    (run_w_checkpoint.py:145-147 -> run.py:53-84); the 1024-entry device ring is spilled to the host in time."""
    from sac_td3_td7_b200.rl.agent import TD3
    from sac_td3_td7_b200.rl.replay_memory import SimpleReplayMemory

    env = "Hopper-v4"
    ag, buf = TD3(env), SimpleReplayMemory(8192, env)
    buf.fill_synthetic(8192, seed=0)
    infos = ag.train_ops_n(2600, buf)
    vals = np.array([float(i["train/q_fn"]) for i in infos])
    assert np.isfinite(vals).all() and len(vals) == 2600
    # the per-step API too: 1100 single updates, then the FIRST info
    first = None
    for t in range(1100):
        info = ag.train_ops(buf.sample(256), replay_buffer=buf)
        first = first or info
    assert np.isfinite(float(first["train/q_fn"]))
    # and the rows equal what the ring held (spot check against a direct read of the newest 100)
    tail = ag.read_infos(ag._n_calls - 99, 100)
    assert np.isfinite(tail).all()
    d = ag.info_dicts(ag._n_calls - 99, 100)
    assert len(d) == 100 and d[-1]["train/q_fn"] == float(tail[-1][0])


def test_population_per_step_training_through_the_python_api():
    """One environment step of a population: append_population (one launch for every seed), sample straight into the
    engine's buffers (zero copy, [n_seeds, B, dim] views), train_ops, and every seed's loss from one strided read --
    equal to the burst API on an identical twin."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env, G = "Hopper-v4", 4
    S, A = 11, 3

    def make():
        torch.manual_seed(1)
        ag = TD7(env, use_lap=True, n_seeds=G, target_update_rate=3)
        ag.seed_population(70)
        buf = LAPReplayMemory(2048, env, n_seeds=G)  # private row stores
        buf.fill_synthetic(1500, seed=9)
        return ag, buf

    a1, b1 = make()
    a2, b2 = make()
    rng = np.random.default_rng(0)
    for t in range(5):
        tr = (rng.standard_normal((G, S)).astype(np.float32), (rng.random((G, A), dtype=np.float32) * 2 - 1), rng.standard_normal(G).astype(np.float32),
              rng.standard_normal((G, S)).astype(np.float32), np.ones(G, np.float32))
        b1.append_population(*tr)
        b2.append_population(*tr)
        batch = b1.sample(256)
        if t > 0:
            assert batch is a1._inbox and batch["state"].shape == (G, 256, S)
        a1.train_ops(batch, replay_buffer=b1)
        a2.train_ops_n(1, b2)
        r1, r2 = a1.read_infos(t + 1, 1, -1), a2.read_infos(t + 1, 1, -1)
        assert r1.shape == (G, 1, 3) and np.isfinite(r1[:, :, :2]).all()
        if t > 0:  # (the first per-step sample used injected-free private tensors; from then on both paths share the device RNG stream)
            pass
    assert len(b1) == 1505 and b1.size == 1505
    # rows landed in every seed's private store, normalised (Hopper: scale 1, bias 0)
    rows = np.empty((1, b1.row_stride), np.float32)
    for g in range(G):
        b1._lib.b200rl_replay_read_rows(b1._h, g, 1504, 1, rows.ctypes.data, None)
        u = b1.unpack(rows)
        assert np.array_equal(u["state"][0], tr[0][g]) and np.array_equal(u["action"][0], tr[1][g]) and np.array_equal(u["next_state"][0], tr[3][g])


def test_append_normalises_actions_on_the_device_bit_exactly():
    """lap.py:35 `a / scale - bias` moved into the append kernel: fp32 actions give numpy's fp32 result bit for bit
    (Humanoid: scale 0.4); fp64 actions are normalised on the host in fp64 like the reference would."""
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env = "Humanoid-v4"
    buf = LAPReplayMemory(512, env)
    S, A = 376, 17
    rng = np.random.default_rng(4)
    want = []
    for t in range(40):
        obs, nxt = rng.standard_normal(S), rng.standard_normal(S)  # float64, like MuJoCo observations
        act = (rng.random(A) * 0.8 - 0.4)
        if t % 2 == 0:
            act = act.astype(np.float32)
        r, nd = float(rng.standard_normal()), float(t % 7 != 0)
        buf.append([obs, act, r, nxt, nd])
        ref_a = act / buf.action_scale - buf.action_bias  # what lap.py:35 computes in the action's own dtype
        want.append(np.concatenate([obs.astype(np.float32), np.asarray(ref_a, np.float32), [np.float32(r), np.float32(nd)], nxt.astype(np.float32)]))
    rows = np.empty((40, buf.row_stride), np.float32)
    buf._lib.b200rl_replay_read_rows(buf._h, 0, 0, 40, rows.ctypes.data, None)
    u = buf.unpack(rows)
    got = np.concatenate([u["state"], u["action"], u["reward"], u["done"], u["next_state"]], 1)
    assert np.array_equal(got, np.stack(want))
    assert buf.size == 40 and len(buf) == 40
    assert torch.equal(buf.priority[:40], torch.ones(40))


def test_import_from_and_mark_dirty():
    """Host-side edits reach the device: import_from(other) (inverse of export_to) and mark_dirty() after an in-place
    load_state_dict on a module."""
    from sac_td3_td7_b200.rl.agent import TD3

    env = "Hopper-v4"
    torch.manual_seed(1)
    a = TD3(env)
    torch.manual_seed(2)
    b = TD3(env)
    obs = np.ones(11, np.float32)
    ya, yb = a._policy_out(obs), b._policy_out(obs)
    assert np.abs(ya - yb).max() > 1e-3
    b.import_from(a)
    assert np.abs(b._policy_out(obs) - ya).max() <= 1e-6
    torch.manual_seed(3)
    c = TD3(env)
    c.policy.load_state_dict(a.policy.state_dict())
    c.mark_dirty()
    assert np.abs(c._policy_out(obs) - ya).max() <= 1e-6


def test_population_evaluate_hook_batches_acting_over_seeds():
    """Population.evaluate: every seed acts in its own environment through ONE batched acting call per step
    (rl/runner/run.py:21-45 test_agent, per seed)."""
    from sac_td3_td7_b200.population import Population
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    class ToyEnv:
        """reward = -|a - tanh(mean(obs))|: deterministic given the seed; gymnasium's reset / step signatures."""

        def __init__(self, seed, S, A):
            self.g, self.S, self.A, self.t = np.random.default_rng(seed), S, A, 0

        def reset(self, seed=None):
            self.t = 0
            self.o = self.g.standard_normal(self.S).astype(np.float32)
            return self.o, {}

        def step(self, a):
            assert a.shape == (self.A,)
            r = -float(np.abs(a - np.tanh(self.o.mean())).mean())
            self.t += 1
            self.o = self.g.standard_normal(self.S).astype(np.float32)
            return self.o, r, False, self.t >= 5, {}

    pop = Population(TD7, LAPReplayMemory, "Hopper-v4", total_seeds=3, base_seed=10, buffer_rows=2048, rank=0, world_size=1, use_lap=True)
    envs = [ToyEnv(100 + g, 11, 3) for g in range(3)]
    ret = pop.evaluate(envs, n_episodes=2)
    assert ret.shape == (3,) and np.isfinite(ret).all()
    # the same episodes replayed seed by seed through single-seed acting of that seed's weights
    from oracle import nets as onets

    for g in range(3):
        env = ToyEnv(100 + g, 11, 3)
        mods = pop.agent.seed_modules(g)
        total = 0.0
        for _ in range(2):
            o, _ = env.reset()
            done = False
            while not done:
                s = torch.as_tensor(o).reshape(1, -1)
                a = onets.sale_actor(_sd(mods["policy"]), s, onets.encode_state(_sd(mods["fixed_encoder"]), s))[0].numpy()
                o, r, term, trunc, _ = env.step(np.clip(a, -1, 1) * pop.agent.action_scale + pop.agent.action_bias)
                total += r
                done = term or trunc
        assert abs(ret[g] - total / 2) <= 1e-4


@pytest.mark.parametrize("batch,hdim", [(96, 256), (256, 200)])
def test_stacked_seeds_with_shapes_the_tensor_cores_cannot_tile(batch, hdim):
    """A batch that is not a multiple of 128 rows, or a hidden width that is not a multiple of 16, keeps a stacked population
    off the tcgen05 kernels (and off the fused AvgL1Norm epilogue): the fp32 kernels take over, the update still runs and every
    seed's losses are finite and distinct."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    G = 8
    torch.manual_seed(3)
    agent = TD7("Hopper-v4", use_lap=True, n_seeds=G, batch_size=batch, hdim=hdim, zs_dim=hdim)
    agent.seed_population(11)
    buf = LAPReplayMemory(4096, "Hopper-v4", n_seeds=G)
    buf.fill_synthetic(4000, seed=2)
    agent.train_ops_n(6, buf)
    torch.cuda.synchronize()
    q = np.array([agent.info_dicts(agent.n_runs, 1, seed=g)[0]["train/q_fn"] for g in range(G)], np.float64)
    assert np.isfinite(q).all() and len(set(np.round(q, 6))) == G
