"""The Python drop-in layer (sac_td3_td7_b200.rl.*) on a real GPU, driven the way the reference's
runner drives it (rl/runner/run.py:87-96), against the oracle."""
import copy
import os
import tempfile

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


def _mk(name):
    from sac_td3_td7_b200.rl.agent import SAC, TD3, TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    algo, env_id, prioritized, kwargs, steps = gu.CASES[name]
    torch.manual_seed(gu.SEED)
    cls = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo]
    agent = cls(env_id, **kwargs)
    buf = (LAPReplayMemory if prioritized else SimpleReplayMemory)(gu.BUFFER_ROWS, env_id)
    buf.fill_synthetic(gu.BUFFER_ROWS - 96, seed=gu.SEED)
    return agent, buf, algo, steps


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td3_hopper_uniform", "sac_walker2d"])
def test_runner_style_loop_matches_reference_golden(name):
    """Same seed => same initial weights as the reference; the reference's recorded draws; the
    reference's recorded infos / indices (free-running, so fp32 drift tolerance on the losses)."""
    g = gu.load(name)
    agent, buf, algo, steps = _mk(name)
    for t in range(steps):
        batch = buf.sample(gu.BATCH, u=torch.tensor(g[f"s{t}/u"]))
        assert np.array_equal(buf.ind, g[f"s{t}/ind"]), f"step {t}"
        assert np.array_equal(batch["action"].cpu().numpy(), g[f"s{t}/batch_action"])
        assert batch["state"].shape == (gu.BATCH, agent.state_dim) and batch["reward"].shape == (gu.BATCH, 1)
        info = agent.train_ops(batch, replay_buffer=buf, noise=gu.step_noise(g, t, algo))
        for k, v in info.items():
            ref = float(g[f"s{t}/info/{k}"])
            if v is None:
                assert np.isnan(ref)
            else:
                assert abs(float(v) - ref) <= 2e-3 * max(abs(ref), 1e-2), (t, k, float(v), ref)
                assert f"{v:.4f}" == f"{float(v):.4f}"
        if buf.prioritized:
            assert abs(buf.max_priority - float(g[f"s{t}/max_priority"])) <= 1e-4 * float(g[f"s{t}/max_priority"])
    assert len(buf) == buf.ptr and buf.size == gu.BUFFER_ROWS - 96
    act = agent.sample(g["act/obs"], True)
    assert np.abs(act - g["act/action"]).max() < 5e-3
    # run.py:47-51 mean over infos with None entries
    vals = [i for i in [info.get("train/q_fn")] if i is not None]
    assert sum(vals) / len(vals) == pytest.approx(float(vals[0]))


def test_burst_api_and_device_rng():
    agent, buf, algo, _ = _mk("td7_halfcheetah_lap")
    infos = agent.train_ops_n(12, buf)
    assert len(infos) == 12 and agent.n_runs == 12
    assert [i["train/policy"] is not None for i in infos] == [(k + 1) % 2 == 0 for k in range(12)]
    q = [float(i["train/q_fn"]) for i in infos]
    assert all(np.isfinite(q)) and all(0.1 < x < 10 for x in q)
    assert buf.max_priority >= 1.0
    pri = buf.priority[: buf.size]
    assert float(pri.min()) >= 1.0 and float(pri.max()) > 1.0  # priorities were written back


def test_state_dict_modules_deepcopy_load_save():
    from sac_td3_td7_b200.rl.agent import TD7, Agent

    agent, buf, algo, _ = _mk("td7_halfcheetah_lap")
    w0 = agent.policy.l1.weight.detach().clone()
    agent.train_ops_n(4, buf)
    w1 = agent.policy.l1.weight.detach().clone()
    assert not torch.equal(w0, w1)                       # host modules follow the device weights
    assert agent.target_policy is agent.policy           # aliased, SURVEY.md B.1
    twin = copy.deepcopy(agent)                           # run_w_checkpoint.py:72
    assert torch.equal(twin.policy.l1.weight, w1) and twin.n_runs == agent.n_runs
    agent.train_ops_n(2, buf)
    assert not torch.equal(agent.q1.q2.weight, twin.q1.q2.weight)
    twin.load_state_dict(agent)                           # run_w_checkpoint.py:140
    assert torch.equal(agent.q1.q2.weight, twin.q1.q2.weight)
    assert torch.equal(agent.fixed_encoder_target.zs1.weight, twin.fixed_encoder_target.zs1.weight)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt.pkl")
        agent.to(torch.device("cpu")).to(torch.device("cuda"))
        agent.save(path)                                  # abc.py:38-47
        back = Agent.load(path)
        assert isinstance(back, TD7) and back.n_runs == agent.n_runs
        assert torch.equal(back.encoder.zsa3.weight, agent.encoder.zsa3.weight)
        obs = np.linspace(-1, 1, agent.state_dim).astype(np.float32)
        assert np.allclose(back.sample(obs, True), agent.sample(obs, True), atol=1e-6)
        # the restored agent keeps training (optimizer moments and counters came along)
        i1 = back.train_ops_n(2, buf)
        assert np.isfinite(float(i1[-1]["train/q_fn"]))


def test_nn_modules_forward_match_oracle():
    from oracle import nets as onets
    from sac_td3_td7_b200.rl import nn as bnn

    torch.manual_seed(3)
    S, A, B = 17, 6, 64
    enc, pol, cri = bnn.SALEEncoder(S, A).cuda(), bnn.SALEActor(S, A).cuda(), bnn.SALECritic(S, A).cuda()
    mlp_pi, mlp_q = bnn.MLPActor(S, 2 * A).cuda(), bnn.MLPCritic(S, A).cuda()
    s, a = torch.randn(B, S), torch.rand(B, A) * 2 - 1
    cpu = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}  # noqa: E731
    zs = enc.encode_state(s.cuda())
    zs_ref = onets.encode_state(cpu(enc), s)
    assert torch.allclose(zs.cpu(), zs_ref, atol=2e-5, rtol=1e-5)
    zsa = enc.encode_state_action(zs, a.cuda())
    assert torch.allclose(zsa.cpu(), onets.encode_state_action(cpu(enc), zs_ref, a), atol=2e-5, rtol=1e-5)
    assert torch.allclose(pol.inference_mean(s.cuda(), zs).cpu(), onets.sale_actor(cpu(pol), s, zs_ref), atol=2e-5)
    q = cri.estimate_q_value(s.cuda(), a.cuda(), zsa, zs)
    assert torch.allclose(q.cpu(), onets.sale_critic(cpu(cri), s, a, zsa.cpu(), zs_ref), atol=2e-5, rtol=1e-5)
    mean, logvar = mlp_pi.inference_mean_logvar(s.cuda())
    ref = onets.mlp(cpu(mlp_pi), s)
    assert torch.allclose(torch.cat([mean, logvar], -1).cpu(), ref, atol=2e-5, rtol=1e-5)
    assert torch.allclose(mlp_q.estimate_q_value(s.cuda(), a.cuda()).cpu(), onets.mlp_critic(cpu(mlp_q), s, a), atol=2e-5, rtol=1e-5)
    x = torch.randn(B, 256)
    assert torch.allclose(bnn.AvgL1Norm(x.cuda()).cpu(), onets.avg_l1_norm(x), atol=1e-6, rtol=1e-5)


def test_append_from_rollout_and_uniform_buffer():
    from oracle import replay as oreplay
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env = "Humanoid-v4"  # action scale 0.4: exercises the a/scale - bias normalisation (lap.py:35)
    buf, ref = LAPReplayMemory(500, env), oreplay.OracleLAP(500, env)
    g = np.random.default_rng(0)
    for i in range(620):
        tr = [g.standard_normal(376).astype(np.float32), g.uniform(-0.4, 0.4, 17).astype(np.float32), float(g.standard_normal()),
              g.standard_normal(376).astype(np.float32), 1.0 - float(i % 40 == 0)]
        buf.append(copy.deepcopy(tr))
        ref.append(copy.deepcopy(tr))
    assert (len(buf), buf.size) == (len(ref), ref.size)
    u = torch.rand(256)
    b, rb = buf.sample(256, u=u), ref.sample(256, u=u)
    assert np.array_equal(buf.ind, ref.ind)
    for k in rb:
        assert torch.equal(b[k].cpu(), rb[k]), k


def test_zero_copy_sample_hand_off():
    """After the first train_ops(batch, replay_buffer=buf) the buffer samples straight into the agent's input
    buffers: sample() returns tensors aliasing them (same dict every call), the results are those of the
    staged path, and a buffer created with zero_copy=False keeps handing out private tensors."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    def run(zero_copy):
        torch.manual_seed(7)
        agent = TD7("Hopper-v4", use_lap=True)
        buf = LAPReplayMemory(4096, "Hopper-v4", zero_copy=zero_copy)
        buf.fill_synthetic(4000, seed=1)
        g = torch.Generator().manual_seed(3)
        out, batches = [], []
        for t in range(6):
            u = torch.rand(256, generator=g)
            nz = torch.randn(256, agent.action_dim, generator=g)
            batch = buf.sample(256, u=u)
            batches.append(batch)
            ind = buf.ind.copy()
            info = agent.train_ops(batch, replay_buffer=buf, noise=nz)
            out.append((ind, batch["reward"].cpu().numpy().copy(), float(info["train/q_fn"]), buf.max_priority))
        return agent, buf, out, batches

    a1, b1, o1, batches1 = run(True)
    a0, b0, o0, batches0 = run(False)
    for x, y in zip(o1, o0):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2] and x[3] == y[3]
    assert batches1[1] is a1._inbox and batches1[2] is batches1[1]          # aliasing, from the second call on
    # (rows of the engine's state buffer start every round_up(state_dim, 4) floats: 16-byte aligned operands)
    assert batches1[1]["state"].stride(0) == (a1.state_dim + 3) // 4 * 4
    assert batches1[1]["next_state"].data_ptr() == batches1[1]["state"].data_ptr() + 4 * 256 * batches1[1]["state"].stride(0)
    assert batches0[1] is not batches0[2] and a0._inbox is None or batches0[1] is not a0._inbox
    # another agent fed from the same buffer takes the staged path and then becomes the sink
    other = TD7("Hopper-v4", use_lap=True)
    batch = b1.sample(256)
    assert batch is a1._inbox
    info = other.train_ops(batch, replay_buffer=b1)
    assert np.isfinite(float(info["train/q_fn"]))
    assert b1.sample(256) is other._inbox


def test_single_kernel_append_from_pinned_and_device_rows():
    """Per-step appends whose source the GPU can read (pinned host memory, device memory) are one kernel (row copy +
    priorities + tree repair); the result must be what the staged path (pageable numpy rows) produces, across the wrap."""
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    for cls in (LAPReplayMemory, SimpleReplayMemory):
        bufs = [cls(300, "Hopper-v4") for _ in range(3)]
        g = np.random.default_rng(9)
        stride = bufs[0].row_stride
        pinned = torch.empty((8, stride), dtype=torch.float32).pin_memory()
        for it in range(120):  # 120 x 3 rows: wraps the 300-row ring
            rows = g.standard_normal((3, stride)).astype(np.float32)
            slot = pinned[(it % 2) * 4:(it % 2) * 4 + 3]
            torch.cuda.synchronize()
            slot.copy_(torch.from_numpy(rows))
            bufs[0].append_rows(rows)                       # pageable: cudaMemcpyAsync + kernel
            bufs[1].append_rows(slot.numpy())               # pinned: one kernel reading host memory
            bufs[2].append_rows(torch.from_numpy(rows).cuda())  # device rows: one kernel
        u = torch.rand(64)
        ref = bufs[0].sample(64, u=u)
        for b in bufs[1:]:
            got = b.sample(64, u=u)
            assert (b.ptr, b.size) == (bufs[0].ptr, bufs[0].size)
            assert np.array_equal(b.ind, bufs[0].ind)
            for k in ref:
                assert torch.equal(got[k], ref[k]), k


def test_info_dicts_after_a_burst_and_export_to_reference_classes():
    """ info_dicts() = the dicts train_ops would have returned, from one ring read; export_to()
    fills an agent of the reference's classes (here: a stand-in with the same attribute / state_dict layout, the real
    classes cannot travel to the GPU box) so that the oracle forward of the exported weights equals agent.sample()."""
    from oracle import nets as onets
    from sac_td3_td7_b200.rl import nn as bnn
    from sac_td3_td7_b200.rl.agent import TD7

    agent, buf, algo, _ = _mk("td7_halfcheetah_lap")
    per_call = [agent.train_ops(buf.sample(gu.BATCH), replay_buffer=buf) for _ in range(3)]
    burst = agent.train_ops_n(5, buf)
    dicts = agent.info_dicts(1, 8)
    assert [d["train/policy"] is None for d in dicts] == [(k + 1) % 2 != 0 for k in range(8)]
    for d, ref in zip(dicts, per_call + burst):
        for k, v in ref.items():
            assert (d[k] is None) == (v is None) and (v is None or d[k] == float(v))

    class RefAgent:  # attribute names and module classes of rl/agent/td7.py:53-66
        def __init__(self, S, A):
            self.policy, self.q1, self.q2 = bnn.SALEActor(S, A), bnn.SALECritic(S, A), bnn.SALECritic(S, A)
            self.encoder, self.fixed_encoder, self.fixed_encoder_target = bnn.SALEEncoder(S, A), bnn.SALEEncoder(S, A), bnn.SALEEncoder(S, A)
            self.target_q1, self.target_q2 = bnn.SALECritic(S, A), bnn.SALECritic(S, A)
            self.target_policy = self.policy

    ref = RefAgent(agent.state_dim, agent.action_dim)
    agent.export_to(ref)
    obs = np.linspace(-1, 1, agent.state_dim).astype(np.float32)
    cpu = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}  # noqa: E731
    s = torch.tensor(obs)[None]
    a = onets.sale_actor(cpu(ref.policy), s, onets.encode_state(cpu(ref.fixed_encoder), s))[0].numpy()
    assert np.abs(agent.sample(obs, True) - (a * agent.action_scale + agent.action_bias)).max() < 1e-5
    assert torch.equal(ref.q1.q2.weight, agent.q1.q2.weight.cpu()) and torch.equal(ref.encoder.zs1.weight, agent.encoder.zs1.weight.cpu())


def test_sum_tree_stays_exact_over_a_long_run():
    """600 TD7 updates (two hard updates with reset_max_priority, appends wrapping the ring in between): the sampler's
    incrementally repaired sum tree must give the same indices as a tree rebuilt from the priorities, and as
    torch.cumsum + searchsorted on the priorities read back (the reference's sampler, lap.py:45-54)."""
    from sac_td3_td7_b200 import _lib
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    torch.manual_seed(11)
    env, cap = "Hopper-v4", 20_000
    agent, buf = TD7(env, use_lap=True), LAPReplayMemory(cap, env)
    buf.fill_synthetic(cap - 50, seed=2)
    g = np.random.default_rng(4)
    for rnd in range(12):
        agent.train_ops_n(50, buf)
        buf.append_rows(g.standard_normal((int(g.integers(1, 40)), buf.row_stride)).astype(np.float32))
    assert agent.n_runs == 600 and buf.size == cap
    u = torch.rand(256, generator=torch.Generator().manual_seed(1))
    buf.zero_copy = False
    buf._sink = None
    a = buf.sample(256, u=u)
    ind_incremental = buf.ind.copy()
    pri = buf.priority
    prefix = torch.cumsum(pri[: buf.size], 0)
    ref_ind = torch.searchsorted(prefix, u * prefix[-1]).numpy()
    assert np.array_equal(ind_incremental, ref_ind)
    _lib.check(buf._lib.b200rl_replay_set_cursor(buf._h, 0, buf.ptr, buf.size))  # invalidates the tree: next sample rebuilds it
    b = buf.sample(256, u=u)
    assert np.array_equal(buf.ind, ind_incremental)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert float(pri.max()) == pytest.approx(buf.max_priority) or buf.max_priority >= float(pri.max())


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td3_hopper_uniform", "sac_walker2d"])
def test_losses_posted_to_the_host_mailbox_equal_the_ring(name):
    """float(info[k]) of the newest update is served from the pinned host mailbox (posted by the kernel that computes
    the value, before the rest of the update has run); it must be bit-identical to the device info ring, for every
    key of every step, and older updates must still be readable from the ring."""
    agent, buf, algo, _ = _mk(name)
    kept = []
    for t in range(6):
        info = agent.train_ops(buf.sample(gu.BATCH), replay_buffer=buf)
        step = agent._n_calls
        got = {k: (None if v is None else float(v)) for k, v in info.items()}
        assert got == agent.info_dicts(step, 1)[0], (t, got)
        kept.append((step, agent._info_dict(step, agent._actor_step(step))))
    for step, lazy in kept[:-1]:  # not the newest any more: ring path
        assert {k: (None if v is None else float(v)) for k, v in lazy.items()} == agent.info_dicts(step, 1)[0]


def test_no_cpu_fallback_paths():
    from sac_td3_td7_b200 import _lib
    from sac_td3_td7_b200.rl import nn as bnn

    with pytest.raises(_lib.B200RLError):
        bnn.AvgL1Norm(torch.randn(4, 8))  # CPU tensor: refused, not silently computed on the host


def test_stacked_population_on_tensor_cores():
    """8 stacked seeds -> the tcgen05 (TF32) path.  Seed 0 shares weights, transitions and Philox
    streams with a single fp32 (FFMA) agent, so its losses must track that agent within the TF32
    tolerance; the other seeds must be distinct, finite learners."""
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    env, G, steps = "Ant-v4", 8, 8
    pop = TD7(env, use_lap=True, n_seeds=G)
    pop.seed_population(500)
    buf = LAPReplayMemory(8192, env, n_seeds=G, share_rows=True)
    buf.fill_synthetic(8192, seed=3)
    torch.manual_seed(500)
    one = TD7(env, use_lap=True)
    b1 = LAPReplayMemory(8192, env)
    b1.fill_synthetic(8192, seed=3)
    pop.train_ops_n(steps, buf)
    one.train_ops_n(steps, b1)
    ref = one.read_infos(1, steps)
    got0 = pop.read_infos(1, steps, seed=0)
    assert np.all(np.isfinite(got0))
    # encoder / critic losses (columns 0, 1) every step; policy loss (column 2) on even steps only
    assert np.abs(got0[:, :2] - ref[:, :2]).max() <= 2e-2 * np.abs(ref[:, :2]).max(), (got0, ref)
    assert np.abs(got0[1::2, 2] - ref[1::2, 2]).max() <= 5e-2 * max(np.abs(ref[1::2, 2]).max(), 0.05), (got0, ref)
    others = [pop.read_infos(1, steps, seed=g) for g in range(1, G)]
    for o in others:
        assert np.all(np.isfinite(o[:, :2])) and not np.allclose(o[:, :2], got0[:, :2])


@pytest.mark.parametrize("algo", ["TD3", "SAC"])
def test_stacked_mlp_population_on_tensor_cores(algo):
    """TD3 / SAC populations (8 stacked seeds, uniform replay, shared rows) on the tcgen05 path: seed 0 shares weights,
    transitions and Philox streams with a single fp32 agent and must track its critic loss within the TF32 tolerance;
    the other seeds are distinct, finite learners."""
    from sac_td3_td7_b200.rl import agent as bagent
    from sac_td3_td7_b200.rl.replay_memory import SimpleReplayMemory

    env, G, steps = "Walker2d-v4", 8, 8
    cls = getattr(bagent, algo)
    pop = cls(env, n_seeds=G)
    pop.seed_population(700)
    buf = SimpleReplayMemory(8192, env, n_seeds=G, share_rows=True)
    buf.fill_synthetic(8192, seed=4)
    torch.manual_seed(700)
    one = cls(env)
    b1 = SimpleReplayMemory(8192, env)
    b1.fill_synthetic(8192, seed=4)
    pop.train_ops_n(steps, buf)
    one.train_ops_n(steps, b1)
    ref = one.read_infos(1, steps)
    got0 = pop.read_infos(1, steps, seed=0)
    assert np.all(np.isfinite(got0[:, 0]))
    assert np.abs(got0[:, 0] - ref[:, 0]).max() <= 2e-2 * np.abs(ref[:, 0]).max(), (got0[:, 0], ref[:, 0])
    for g in range(1, G):
        o = pop.read_infos(1, steps, seed=g)
        assert np.all(np.isfinite(o[:, 0])) and not np.allclose(o[:, 0], got0[:, 0])
