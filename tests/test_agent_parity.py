"""TD7 / TD3 / SAC update steps through the C ABI vs the oracle, on the reference-recorded cases.

Each step is checked from IDENTICAL state (weights, Adam moments and priorities are copied from
the oracle before the step -- "teacher forcing"), because k-step trajectories of any two fp32
implementations drift through Adam; a free-running run is checked separately
with a drift bound.  Per step, with the draws the reference itself made (golden uniforms/noise):
  * sampled indices: bit-exact        * batch: bit-exact
  * losses / info: <= 1e-5 relative   * q target, priorities: <= 1e-5 normwise
  * gradients: <= 1e-5 normwise (max|d|/max|ref| per tensor)
  * updated weights: <= 1e-5 normwise; target nets after hard copy / Polyak likewise
The oracle itself is pinned bit-exact to the reference by tests/test_oracle_golden.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import devlib
import golden_util as gu
from sac_td3_td7_b200 import _lib

pytestmark = pytest.mark.usefixtures("gemm_path")  # both product loops of the small-tile kernel (conftest.py)

TD7_NETS = [("encoder", devlib.SALE_ENCODER), ("policy", devlib.SALE_ACTOR), ("q1", devlib.SALE_CRITIC), ("q2", devlib.SALE_CRITIC),
            ("target_q1", devlib.SALE_CRITIC), ("target_q2", devlib.SALE_CRITIC), ("fixed_encoder", devlib.SALE_ENCODER),
            ("fixed_encoder_target", devlib.SALE_ENCODER)]
MLP_NETS = [("policy", devlib.MLP), ("q1", devlib.MLP), ("q2", devlib.MLP), ("target_q1", devlib.MLP), ("target_q2", devlib.MLP)]
TRAINED = {"encoder": "optim_encoder", "policy": "optim_policy", "q1": "optim_q_fns", "q2": "optim_q_fns"}


def nerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# Layers whose output feeds AvgL1Norm.  d/dx |x| = sign(x) is discontinuous at 0: when a pre-norm
# activation is within rounding of zero (|x| ~ 1e-8 happens about once per 10^5 rows) two correct fp32
# implementations disagree on its sign and the gradient of THAT output unit (one row of W, one bias
# entry) differs -- a reference-inherent branch flip.  For these layers only, up to
# MAX_FLIP_ROWS output rows per tensor may be excluded before the strict tolerance applies; an
# indexing or arithmetic bug would disturb far more rows than that.
NORM_FED = {("q1", "q01"), ("q2", "q01"), ("policy", "l0"), ("encoder", "zs3"),
            # ... and the hard copies of those tensors taken at a target sync (td7.py:279-285)
            ("target_q1", "q01"), ("target_q2", "q01"), ("fixed_encoder", "zs3"), ("fixed_encoder_target", "zs3")}
# The same holds for layers whose output passes through ReLU (the MLP nets of TD3 / SAC, l1 / l2 of the SALE actor): relu'
# jumps at 0, so a pre-activation within rounding of zero (seen once in the goldens: unit 216 of q1.mlp.0, td3_ant_lap
# step 3, on the 3xTF32 product loop -- the FFMA loop happens to round it to the oracle's side) switches one
# (sample, unit) term of that unit's weight-row and bias gradient on or off.
RELU_FED = {("q1", "mlp.0"), ("q1", "mlp.2"), ("q2", "mlp.0"), ("q2", "mlp.2"), ("policy", "mlp.0"), ("policy", "mlp.2"), ("policy", "l1"), ("policy", "l2")}
MAX_FLIP_ROWS = 2


def nerr_rows(arr, ref, net, layer, algo):
    a, b = np.asarray(arr, np.float64), np.asarray(ref, np.float64)
    d = np.abs(a - b).reshape(a.shape[0], -1).max(1) / max(np.abs(b).max(), 1e-30 if b.ndim == 0 else 1e-12)
    if (((algo == "TD7" and (net, layer) in NORM_FED) or (net, layer) in RELU_FED)) and d.size > MAX_FLIP_ROWS:
        d = np.sort(d)[: d.size - MAX_FLIP_ROWS]
    return d.max(initial=0.0)


def make_c_agent(dev, algo, oracle, kwargs, keep_grads=1, groups=1):
    h = C.c_void_p()
    S, A = oracle.state_dim, oracle.action_dim
    if algo == "TD7":
        cfg = _lib.TD7Config(state_dim=S, action_dim=A, hdim=256, zs_dim=256, batch=gu.BATCH, groups=groups, discount=0.99, policy_lr=3e-4, critic_lr=3e-4,
                             target_update_rate=kwargs.get("target_update_rate", 250), policy_freq=2, target_policy_noise=0.2, noise_clip=0.5,
                             use_lap=int(kwargs.get("use_lap", False)), keep_grads=keep_grads)
        dev.check(dev.lib.b200rl_td7_create(C.byref(cfg), C.byref(h)))
    else:
        cfg = _lib.MlpConfig(algo=0 if algo == "TD3" else 1, state_dim=S, action_dim=A, hdim=256, batch=gu.BATCH, groups=groups, discount=0.99, policy_lr=3e-4,
                             critic_lr=3e-4, tau=0.005, policy_freq=2, target_policy_noise=0.2, noise_clip=0.5, min_log_std=-20.0, max_log_std=2.0, tmp=-1.0,
                             use_lap=int(kwargs.get("use_lap", False)), keep_grads=keep_grads)
        dev.check(dev.lib.b200rl_mlp_create(C.byref(cfg), C.byref(h)))
    return h.value


def net_table(algo):
    return TD7_NETS if algo == "TD7" else MLP_NETS


def push_state(dev, h, algo, oracle, group=0):
    """oracle -> device: parameters and Adam moments of every network (+ SAC temperature)."""
    onets = oracle.nets()
    for nid, (name, layers) in enumerate(net_table(algo)):
        dev.write_net(h, nid, devlib.params_to_tensors(onets[name], layers), which=0, group=group)
        if name in TRAINED:
            opt = getattr(oracle, TRAINED[name])
            for which, key in ((1, "exp_avg"), (2, "exp_avg_sq")):
                ts = []
                for ln in layers:
                    w, b = onets[name][ln + ".weight"], onets[name][ln + ".bias"]
                    sw, sb = opt.state.get(w, {}), opt.state.get(b, {})
                    ts.append((sw[key].numpy() if key in sw else np.zeros(tuple(w.shape), np.float32),
                               sb[key].numpy() if key in sb else np.zeros(tuple(b.shape), np.float32)))
                dev.write_net(h, nid, ts, which=which, group=group)
    if algo == "SAC":
        st = oracle.optim_tmp.state.get(oracle.tmp, {})
        vals = [oracle.tmp.detach().numpy(), st["exp_avg"].numpy() if st else np.zeros(1, np.float32), st["exp_avg_sq"].numpy() if st else np.zeros(1, np.float32)]
        for which, v in enumerate(vals):
            v = np.ascontiguousarray(v, np.float32)
            dev.check(dev.lib.b200rl_agent_write(h, 5, which, group, v.ctypes.data, None))


def compare_nets(dev, h, algo, oracle, tol, what="", lr=3e-4):
    """Per-tensor normwise check of every network against the oracle.

    Adam turns rounding noise on a near-zero gradient into a visible fraction of an lr-sized step
    (update = lr*g/(|g|+1e-8) at t=1), so for tensors updated this step the
    strict tolerance applies to elements whose reference gradient is >= 1e-6 (100x Adam's eps) and
    the few elements below that are only required to stay within one Adam step (2*lr)."""
    onets = oracle.nets()
    worst = 0.0
    for nid, (name, layers) in enumerate(net_table(algo)):
        got = dev.read_net(h, nid)
        grads = oracle.last.get("grad_" + name) if name in TRAINED else None
        for (w, b), ln in zip(got, layers):
            for arr, key in ((w, ".weight"), (b, ".bias")):
                ref = onets[name][ln + key].detach().numpy()
                gref = grads.get(ln + key) if grads else None
                if gref is not None:
                    strong = np.abs(gref.numpy()) >= 1e-6
                    assert np.abs(arr - ref)[~strong].max(initial=0.0) <= 2 * lr, f"{what} {name}.{ln}{key}: weak-gradient element moved more than an Adam step"
                    assert strong.mean() > 0.5 or not strong.size, f"{what} {name}.{ln}{key}: mask too aggressive"
                    e = nerr_rows(np.where(strong, arr, ref), ref, name, ln, algo)
                elif algo == "TD7" and (name, ln) in NORM_FED and arr.ndim == 2:
                    e = nerr_rows(arr, ref, name, ln, algo)  # a hard copy of a trained tensor carries its flip rows along
                else:
                    # zero-initialised MLP biases are O(lr) after the first steps: floor the scale at 1e-3
                    e = np.abs(arr.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-3)
                worst = max(worst, e)
                assert e <= tol, f"{what} {name}.{ln}{key}: normwise error {e:.3e} > {tol}"
    return worst


def compare_grads(dev, h, algo, oracle, tol):
    for nid, (name, layers) in enumerate(net_table(algo)):
        key = "grad_" + name
        if name not in TRAINED or key not in oracle.last:
            continue
        got = dev.read_net(h, nid, which=3)
        for (w, b), ln in zip(got, layers):
            for arr, suffix in ((w, ".weight"), (b, ".bias")):
                ref = oracle.last[key][ln + suffix]
                if ref is None:
                    continue
                e = nerr_rows(arr, ref.numpy(), name, ln, algo)
                assert e <= tol, f"grad {name}.{ln}{suffix}: normwise error {e:.3e} > {tol}"


INFO_COLS = {"TD7": ["train/encoder", "train/q_fn", "train/policy"], "TD3": ["train/q_fn", "train/policy", "norm/policy"],
             "SAC": ["train/q_fn", "tmp", "norm/tmp", "train/policy", "train/tmp", "entropy"]}


def setup_case(dev, name, keep_grads=1):
    algo, env_id, prioritized, kwargs, steps = gu.CASES[name]
    g = gu.load(name)
    oracle, obuf, _, _ = gu.make_oracle(name)
    S, A = oracle.state_dim, oracle.action_dim
    h = make_c_agent(dev, algo, oracle, kwargs, keep_grads)
    r = dev.replay_create(gu.BUFFER_ROWS, S, A, prioritized=prioritized)
    n = gu.BUFFER_ROWS - 96
    rows = dev.pack_rows(r, obuf.state[:n].astype(np.float32), obuf.action[:n].astype(np.float32), obuf.reward[:n], obuf.next_state[:n].astype(np.float32), obuf.float_done[:n])
    dev.replay_append(r, rows)
    return algo, g, oracle, obuf, h, r, S, A, steps, prioritized


def c_step(dev, h, r, S, A, u, noise, algo):
    idx, batch, ptrs = dev.replay_sample(r, gu.BATCH, S, A, u=u)
    dev.check(dev.lib.b200rl_agent_stage_batch(h, 0, ptrs["state"], ptrs["action"], ptrs["reward"], ptrs["next_state"], ptrs["done"], None))
    if algo == "SAC":
        nz = np.stack([noise[0].numpy(), noise[1].numpy()])
    else:
        nz = noise.numpy()
    dev.check(dev.lib.b200rl_agent_update(h, r, dev.to_dev(nz), None))
    dev.check(dev.lib.b200rl_stream_sync(None))
    return idx[0], batch


@pytest.mark.parametrize("name", list(gu.CASES))
def test_step_parity_teacher_forced(dev, name):
    torch.set_num_threads(1)
    algo, g, oracle, obuf, h, r, S, A, steps, prioritized = setup_case(dev, name)
    if dev.backend == "emu" and name == "td7_humanoid_lap":
        steps = 2
    for t in range(steps):
        push_state(dev, h, algo, oracle)
        if prioritized:
            dev.replay_write_priority(r, obuf.priority[: obuf.size].numpy(), max_priority=float(obuf.max_priority))
        u, noise = g[f"s{t}/u"], gu.step_noise(g, t, algo)
        obatch = obuf.sample(gu.BATCH, u=torch.tensor(u))
        oinfo = oracle.train_ops(obatch, replay_buffer=obuf, noise=noise)
        idx, batch = c_step(dev, h, r, S, A, u, noise, algo)
        # sampling: bit exact against the reference's recorded indices and the oracle batch
        assert np.array_equal(idx, g[f"s{t}/ind"]), f"step {t}: sampled indices differ"
        for k in ("state", "action", "reward", "next_state", "done"):
            assert np.array_equal(batch[k][0], obatch[k].numpy()), f"step {t}: batch[{k}] differs"
        # info
        info = dev.read_info(h, t + 1)
        for col, key in enumerate(INFO_COLS[algo]):
            if oinfo.get(key) is None:
                continue
            ref = oinfo[key]
            assert abs(info[col] - ref) <= 1e-5 * max(abs(ref), 1e-3), f"step {t}: {key} {info[col]} vs {ref}"
            assert abs(ref - float(g[f"s{t}/info/{key}"])) <= 1e-6 * max(abs(ref), 1e-3)
        assert nerr(dev.debug(h, 0, gu.BATCH), oracle.last["q_target"].numpy().reshape(-1)) <= 1e-5
        assert nerr(dev.debug(h, 2, gu.BATCH), oracle.last["q1"].numpy().reshape(-1)) <= 1e-5
        if "pi_action" in oracle.last:
            assert np.abs(dev.debug(h, 4, gu.BATCH * A) - oracle.last["pi_action"].numpy().reshape(-1)).max() <= 2e-6
        if prioritized and "priority" in oracle.last:
            assert nerr(dev.debug(h, 1, gu.BATCH), oracle.last["priority"].numpy()) <= 1e-5
            pri = dev.replay_priority(r, obuf.size)
            assert nerr(pri, obuf.priority[: obuf.size].numpy()) <= 1e-5
            assert abs(dev.replay_state(r)[2] - float(obuf.max_priority)) <= 1e-5 * float(obuf.max_priority)
        if algo == "TD7":
            rng4 = np.empty(4, np.float32)
            dev.check(dev.lib.b200rl_td7_value_range(h, 0, rng4.ctypes.data, None))
            ref4 = np.array([oracle.value_max, oracle.value_min, oracle.value_target_max, oracle.value_target_min])
            assert np.abs(rng4 - ref4).max() <= 1e-5 * max(np.abs(ref4).max(), 1.0)
            # value range is state too: force it
            dev.check(dev.lib.b200rl_td7_set_value_range(h, 0, ref4.astype(np.float32).ctypes.data, None))
        compare_grads(dev, h, algo, oracle, 1e-5)
        compare_nets(dev, h, algo, oracle, 1e-5, what=f"step {t}")
        if algo == "SAC":
            la = np.empty(1, np.float32)
            dev.check(dev.lib.b200rl_agent_read(h, 5, 0, 0, la.ctypes.data, None))
            assert abs(la[0] - float(oracle.tmp.detach())) <= 1e-6
    dev.check(dev.lib.b200rl_agent_destroy(h))
    dev.check(dev.lib.b200rl_replay_destroy(r))


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td3_hopper_uniform", "sac_walker2d"])
def test_free_running_drift_is_bounded(dev, name):
    """No forcing after initialisation: k updates in a row (priorities feed the next sample).
    Drift is reported and bounded loosely -- this guards state handling across steps (counters,
    Adam bias correction, hard copies, value clipping, ring cursor), not rounding."""
    torch.set_num_threads(1)
    algo, g, oracle, obuf, h, r, S, A, steps, prioritized = setup_case(dev, name, keep_grads=0)
    push_state(dev, h, algo, oracle)
    for t in range(steps):
        u, noise = g[f"s{t}/u"], gu.step_noise(g, t, algo)
        obatch = obuf.sample(gu.BATCH, u=torch.tensor(u))
        oinfo = oracle.train_ops(obatch, replay_buffer=obuf, noise=noise)
        idx, _ = c_step(dev, h, r, S, A, u, noise, algo)
        assert np.array_equal(idx, g[f"s{t}/ind"]), f"step {t}: indices diverged"
        info = dev.read_info(h, t + 1)
        for col, key in enumerate(INFO_COLS[algo]):
            if oinfo.get(key) is not None:
                assert abs(info[col] - oinfo[key]) <= 2e-3 * max(abs(oinfo[key]), 1e-2), f"step {t}: {key} {info[col]} vs {oinfo[key]}"
    worst = compare_nets(dev, h, algo, oracle, 2e-3, what="free-running")
    print(f"{name}: free-running weight drift after {steps} updates = {worst:.2e} (normwise)")
    dev.check(dev.lib.b200rl_agent_destroy(h))
    dev.check(dev.lib.b200rl_replay_destroy(r))


def test_stacked_seeds_equal_independent_agents(dev, monkeypatch):
    """n_seeds = 2 stacked in one engine (gridDim.y = seed) must reproduce, bit for bit, two
    independent single-seed engines fed the same per-seed weights, batches, uniforms and noise.
    (Split-K is pinned to 1: the split factor is chosen from the CTA count, which differs between
    1 and 2 seeds, and changes the fp32 summation order in the last bit.)"""
    monkeypatch.setenv("B200RL_SPLITK", "1")
    from oracle import agents as oagents
    from oracle import replay as oreplay

    torch.set_num_threads(1)
    env, G, steps, rows_n = "Hopper-v4", 2, 4, 2048
    S, A = 11, 3
    kwargs = dict(use_lap=True, target_update_rate=3)
    oracles = []
    for g in range(G):
        torch.manual_seed(100 + g)
        oracles.append(oagents.OracleTD7(env, **kwargs))
    stacked = make_c_agent(dev, "TD7", oracles[0], kwargs, keep_grads=0, groups=G)
    singles = [make_c_agent(dev, "TD7", oracles[g], kwargs, keep_grads=0) for g in range(G)]
    rs = dev.replay_create(rows_n, S, A, groups=G)
    rsingle = [dev.replay_create(rows_n, S, A) for _ in range(G)]
    for g in range(G):
        push_state(dev, stacked, "TD7", oracles[g], group=g)
        push_state(dev, singles[g], "TD7", oracles[g])
        rows = oreplay.synthetic_rows(rows_n, S, A, seed=50 + g)
        packed = dev.pack_rows(rs, rows["state"], rows["action"], rows["reward"], rows["next_state"], rows["done"])
        dev.replay_append(rs, packed, group=g)
        dev.replay_append(rsingle[g], packed)
    rng = np.random.default_rng(0)
    for t in range(steps):
        u = rng.random((G, gu.BATCH), dtype=np.float32)
        nz = rng.standard_normal((G, gu.BATCH, A)).astype(np.float32)
        idx, batch, ptrs = dev.replay_sample(rs, gu.BATCH, S, A, u=u, groups=G)
        for g in range(G):
            off = lambda key, d: ptrs[key] + 4 * g * gu.BATCH * d  # noqa: E731
            dev.check(dev.lib.b200rl_agent_stage_batch(stacked, g, off("state", S), off("action", A), off("reward", 1), off("next_state", S), off("done", 1), None))
        dev.check(dev.lib.b200rl_agent_update(stacked, rs, dev.to_dev(nz), None))
        dev.check(dev.lib.b200rl_stream_sync(None))
        for g in range(G):
            i1, b1, p1 = dev.replay_sample(rsingle[g], gu.BATCH, S, A, u=u[g])
            assert np.array_equal(i1[0], idx[g])
            dev.check(dev.lib.b200rl_agent_stage_batch(singles[g], 0, p1["state"], p1["action"], p1["reward"], p1["next_state"], p1["done"], None))
            dev.check(dev.lib.b200rl_agent_update(singles[g], rsingle[g], dev.to_dev(nz[g]), None))
            dev.check(dev.lib.b200rl_stream_sync(None))
            assert np.array_equal(dev.read_info(stacked, t + 1, group=g), dev.read_info(singles[g], t + 1)), (t, g)
    for g in range(G):
        for nid in range(8):
            for (w1, b1), (w2, b2) in zip(dev.read_net(stacked, nid, group=g), dev.read_net(singles[g], nid)):
                assert np.array_equal(w1, w2) and np.array_equal(b1, b2), (g, nid)
        assert np.array_equal(dev.replay_priority(rs, rows_n, group=g), dev.replay_priority(rsingle[g], rows_n))
        assert dev.replay_state(rs, g)[2] == dev.replay_state(rsingle[g])[2]
    # seeds really are different agents
    assert not np.array_equal(dev.read_net(stacked, 1, group=0)[1][0], dev.read_net(stacked, 1, group=1)[1][0])


# ---------------------------------------------------------------------------------------------------------------------
# The stacked-seed path (configs 4-5): >= 8 seeds run their 128-row-tileable layers on tcgen05 with TF32 inputs (fp32
# operands truncated to 10 mantissa bits by the tensor core, fp32 accumulation in TMEM), the narrow layers on the small-tile
# kernel.  Same teacher-forced protocol as above, every stacked seed loaded with the oracle's state, seeds 0 and G-1
# compared.  STATED TOLERANCE of this path.  Derivation: one TF32 layer is within ~2e-3 normwise of fp32 (operand
# truncation <= 2^-10 relative, averaged over a 256..768-long reduction); a forward chain of 4-7 layers and as many
# backward layers compound it, and every activation-derivative decision taken on a TF32-perturbed pre-activation near zero
# (ReLU: on/off; AvgL1Norm: sign) switches one (sample, unit) term of a gradient row -- frequent at 1e-3 resolution, so the
# max-norm bound of ReLU networks is looser than their Frobenius bound.
#   quantity                                          bound held            measured on B200 (worst over cases / steps / seeds)
#   sampled indices, batch .......................... bit-exact             (the sampler never touches TF32)
#   losses / info ................................... <= 5e-3 relative      2.1e-3
#   q target, q values, priorities .................. <= 5e-3 normwise      1.8e-3 / 4.0e-4
#   gradient tensors, SALE / ELU nets (TD7) ......... <= 5e-2 max-norm      2.5e-2      <= 4e-2 Frobenius   2.3e-2
#   gradient tensors, ReLU MLP nets (TD3 / SAC) ..... <= 2.5e-1 max-norm    1.8e-1      <= 6e-2 Frobenius   4.0e-2
#   updated weights ................................. every element within one Adam step (2 lr) of the oracle's (2.0 lr: a
#                                                     near-zero gradient whose sign differs), rms error of the update
#                                                     <= 0.25 lr (0.18 lr)
#   Adam moments are driven by the gradients (covered by the gradient bounds).
TC_TOL = dict(info=5e-3, q=5e-3, grad_max={"TD7": 5e-2, "TD3": 2.5e-1, "SAC": 2.5e-1}, grad_fro={"TD7": 4e-2, "TD3": 6e-2, "SAC": 6e-2}, step_rms=0.25)


def ferr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30)


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td7_humanoid_lap", "td3_ant_lap", "sac_walker2d"])
def test_step_parity_teacher_forced_on_the_tcgen05_path(dev, name, capsys):
    if dev.backend != "cuda":
        pytest.skip("tcgen05 needs the GPU")
    torch.set_num_threads(1)
    G = 8
    algo, env_id, prioritized, kwargs, steps = gu.CASES[name]
    g = gu.load(name)
    oracle, obuf, _, _ = gu.make_oracle(name)
    S, A = oracle.state_dim, oracle.action_dim
    h = make_c_agent(dev, algo, oracle, kwargs, keep_grads=1, groups=G)
    r = dev.replay_create(gu.BUFFER_ROWS, S, A, groups=G, prioritized=prioritized)
    n = gu.BUFFER_ROWS - 96
    rows = dev.pack_rows(r, obuf.state[:n].astype(np.float32), obuf.action[:n].astype(np.float32), obuf.reward[:n], obuf.next_state[:n].astype(np.float32), obuf.float_done[:n])
    for gg in range(G):
        dev.replay_append(r, rows, group=gg)
    worst = dict(info=(0.0, ""), q=(0.0, ""), prio=(0.0, ""), grad_max=(0.0, ""), grad_fro=(0.0, ""), step=(0.0, ""), elem=(0.0, ""))

    def rec(key, e, what):
        if e > worst[key][0]:
            worst[key] = (float(e), what)

    lr = 3e-4
    for t in range(min(steps, 4)):
        for gg in range(G):
            push_state(dev, h, algo, oracle, group=gg)
            if prioritized:
                dev.replay_write_priority(r, obuf.priority[: obuf.size].numpy(), max_priority=float(obuf.max_priority), group=gg)
        u, noise = g[f"s{t}/u"], gu.step_noise(g, t, algo)
        obatch = obuf.sample(gu.BATCH, u=torch.tensor(u))
        oinfo = oracle.train_ops(obatch, replay_buffer=obuf, noise=noise)
        idx, batch, ptrs = dev.replay_sample(r, gu.BATCH, S, A, u=np.tile(u, (G, 1)), groups=G)
        for gg in range(G):
            off = lambda key, d: ptrs[key] + 4 * gg * gu.BATCH * d  # noqa: E731
            dev.check(dev.lib.b200rl_agent_stage_batch(h, gg, off("state", S), off("action", A), off("reward", 1), off("next_state", S), off("done", 1), None))
        if algo == "SAC":
            nz = np.stack([np.tile(noise[0].numpy(), (G, 1, 1)), np.tile(noise[1].numpy(), (G, 1, 1))])
        else:
            nz = np.tile(noise.numpy(), (G, 1, 1))
        dev.check(dev.lib.b200rl_agent_update(h, r, dev.to_dev(nz), None))
        dev.check(dev.lib.b200rl_stream_sync(None))
        for gg in (0, G - 1):
            tag = f"step {t} seed {gg}"
            assert np.array_equal(idx[gg], g[f"s{t}/ind"]), f"{tag}: sampled indices differ"
            for k in ("state", "action", "reward", "next_state", "done"):
                assert np.array_equal(batch[k][gg], obatch[k].numpy())
            info = dev.read_info(h, t + 1, group=gg)
            for col, key in enumerate(INFO_COLS[algo]):
                ref = oinfo.get(key)
                if ref is None or key in ("norm/policy", "norm/tmp"):
                    continue
                rec("info", abs(info[col] - ref) / max(abs(ref), 1e-2), f"{tag} {key}")
            for which, key in ((0, "q_target"), (2, "q1")):
                rec("q", nerr(dev.debug(h, which, gu.BATCH, group=gg), oracle.last[key].numpy().reshape(-1)), f"{tag} {key}")
            if prioritized and "priority" in oracle.last:
                rec("prio", nerr(dev.debug(h, 1, gu.BATCH, group=gg), oracle.last["priority"].numpy()), f"{tag} priority")
                rec("prio", nerr(dev.replay_priority(r, obuf.size, group=gg), obuf.priority[: obuf.size].numpy()), f"{tag} priority array")
            onets = oracle.nets()
            for nid, (nm, layers) in enumerate(net_table(algo)):
                grads = oracle.last.get("grad_" + nm) if nm in TRAINED else None
                if not grads:
                    continue
                got_g = dev.read_net(h, nid, which=3, group=gg)
                got_p = dev.read_net(h, nid, group=gg)
                for li, ln in enumerate(layers):
                    for ai, suffix in enumerate((".weight", ".bias")):
                        if grads.get(ln + suffix) is None:
                            continue
                        rec("grad_max", nerr(got_g[li][ai], grads[ln + suffix].numpy()), f"{tag} grad {nm}.{ln}{suffix}")
                        rec("grad_fro", ferr(got_g[li][ai], grads[ln + suffix].numpy()), f"{tag} grad {nm}.{ln}{suffix}")
                        d = np.abs(got_p[li][ai].astype(np.float64) - onets[nm][ln + suffix].detach().numpy())
                        rec("elem", d.max() / lr, f"{tag} {nm}.{ln}{suffix}")
                        rec("step", np.sqrt(np.mean(d ** 2)) / lr, f"{tag} {nm}.{ln}{suffix}")
        if algo == "TD7":
            ref4 = np.array([oracle.value_max, oracle.value_min, oracle.value_target_max, oracle.value_target_min], np.float32)
            for gg in range(G):
                dev.check(dev.lib.b200rl_td7_set_value_range(h, gg, ref4.ctypes.data, None))
    with capsys.disabled():
        print(f"\n[tcgen05/TF32 parity {name}] " + " | ".join(f"{k} {v[0]:.3e} ({v[1]})" for k, v in worst.items()))
    dev.check(dev.lib.b200rl_agent_destroy(h))
    dev.check(dev.lib.b200rl_replay_destroy(r))
    assert worst["info"][0] <= TC_TOL["info"], worst["info"]
    assert worst["q"][0] <= TC_TOL["q"], worst["q"]
    assert worst["prio"][0] <= TC_TOL["q"], worst["prio"]
    assert worst["grad_max"][0] <= TC_TOL["grad_max"][algo], worst["grad_max"]
    assert worst["grad_fro"][0] <= TC_TOL["grad_fro"][algo], worst["grad_fro"]
    assert worst["elem"][0] <= 2.002, worst["elem"]
    assert worst["step"][0] <= TC_TOL["step_rms"], worst["step"]


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td7_humanoid_lap"])
def test_step_parity_teacher_forced_through_the_dataflow_kernels(dev, name):
    """The optional persistent dataflow engine (csrc/mega.cuh, b200rl_set_mega): same teacher-forced protocol, same 1e-5
    bars -- its tile order is derived from memory-rectangle overlap, so any missed dependency shows up here."""
    if dev.backend != "cuda":
        pytest.skip("the dataflow kernels are not emulated")
    dev.lib.b200rl_set_mega(1)
    try:
        test_step_parity_teacher_forced(dev, name)
    finally:
        dev.lib.b200rl_set_mega(-1)


@pytest.mark.parametrize("name", ["td7_halfcheetah_lap", "td3_hopper_uniform", "sac_walker2d"])
def test_packed_layout_keeps_state_and_action_blocks_16_byte_aligned(dev, name):
    """A critic's first layer stores [state | pad to 4 | action | pad to 4] (b200rl_agent_layer_split), every other layer one
    dense block; the engine's state / action input buffers have row strides rounded up to 4 floats (b200rl_agent_io_ld);
    pack / unpack through the logical column map is the identity, and the padding columns hold zeros."""
    import ctypes as C

    algo, env_id, prioritized, kwargs, steps = gu.CASES[name]
    oracle, _, _, _ = gu.make_oracle(name)
    S, A = oracle.state_dim, oracle.action_dim
    h = make_c_agent(dev, algo, oracle, kwargs, 0)
    lds, lda = C.c_int(), C.c_int()
    dev.check(dev.lib.b200rl_agent_io_ld(h, C.byref(lds), C.byref(lda)))
    assert lds.value == (S + 3) // 4 * 4 and lda.value == (A + 3) // 4 * 4
    splits = 0
    for nid, (nm, layers) in enumerate(net_table(algo)):
        numel, lay = dev.net_layout(h, nid)
        for li, (wo, bo, o, n, ld, cols) in enumerate(lay):
            sp = C.c_int()
            dev.check(dev.lib.b200rl_agent_layer_split(h, nid, li, C.byref(sp)))
            is_critic_first = nm in ("q1", "q2", "target_q1", "target_q2", "q1_target", "q2_target") and li == 0
            if is_critic_first and S % 4:
                assert sp.value == S and ld == (S + 3) // 4 * 4 + (A + 3) // 4 * 4 and n == S + A
                splits += 1
            else:
                assert sp.value == 0 and ld == (n + 3) // 4 * 4
            assert wo % 4 == 0 and bo % 4 == 0
        got = dev.read_net(h, nid)          # logical [out, in] tensors
        blob = dev.pack_net(h, nid, got)    # back into the padded blob
        again = dev.unpack_net(h, nid, blob)
        for (w0, b0), (w1, b1) in zip(got, again):
            assert np.array_equal(w0, w1) and np.array_equal(b0, b1)
        for (wo, bo, o, n, ld, cols) in lay:  # padding columns are zero
            W = blob[wo:wo + o * ld].reshape(o, ld)
            mask = np.ones(ld, bool)
            mask[cols] = False
            assert np.all(W[:, mask] == 0)
    if S % 4:
        assert splits >= 2
    dev.check(dev.lib.b200rl_agent_destroy(h))
