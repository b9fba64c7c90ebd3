"""Pin the oracle against vectors recorded from the unmodified reference (CPU)."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import replay as oreplay


def _close(a, b, rtol=0.0, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", list(gu.CASES))
def test_agent_trajectory_matches_reference(name):
    torch.set_num_threads(1)
    g = gu.load(name)
    agent, buf, algo, steps = gu.make_oracle(name)
    # initial weights regenerate identically
    for mod_name, params in agent.nets().items():
        for k, v in params.items():
            _close(gu.checksum(v), g[f"init/{mod_name}/{k}"])
    for t in range(steps):
        u = torch.tensor(g[f"s{t}/u"])
        batch = buf.sample(gu.BATCH, u=u)
        assert np.array_equal(buf.ind, g[f"s{t}/ind"])
        _close(gu.checksum(batch["state"]), g[f"s{t}/batch_state_ck"])
        _close(batch["action"].numpy(), g[f"s{t}/batch_action"])
        info = agent.train_ops(batch, replay_buffer=buf, noise=gu.step_noise(g, t, algo))
        for k, v in info.items():
            ref = float(g[f"s{t}/info/{k}"])
            if v is None:
                assert np.isnan(ref), k
            else:
                _close(v, ref, rtol=1e-7)
        if buf.prioritized:
            _close(buf.priority[buf.ind].numpy(), g[f"s{t}/priority_at_ind"])
            _close(float(buf.max_priority), float(g[f"s{t}/max_priority"]))
            _close(gu.checksum(buf.priority[: buf.size]), g[f"s{t}/priority_ck"])
        if algo == "TD7":
            _close([agent.value_max, agent.value_min, agent.value_target_max, agent.value_target_min], g[f"s{t}/value_range"])
        if algo == "SAC":
            _close(float(agent.tmp.detach()), float(g[f"s{t}/tmp"]))
        for mod_name, params in agent.nets().items():
            for k, v in params.items():
                _close(gu.checksum(v), g[f"s{t}/w/{mod_name}/{k}"], rtol=1e-12)
    _close(agent.sample(g["act/obs"], True), g["act/action"], rtol=1e-7)


def test_replay_sampling_cases():
    g = gu.load("replay_cases")
    cap = 3000
    for nm in ("ones", "rand", "zeros_mixed", "pow"):
        for size in (cap, 1777, 1):
            key = f"samp/{nm}/{size}"
            pri, u, ind = g[key + "/priority"], g[key + "/u"], g[key + "/ind"]
            # numpy restatement
            assert np.array_equal(oreplay.sample_indices_numpy(pri[:size], u), ind), key
            # torch restatement through the oracle class
            buf = oreplay.OracleLAP(cap, "Hopper-v4")
            buf.priority[:] = torch.tensor(pri)
            buf.size = size
            buf.sample(256, u=torch.tensor(u))
            assert np.array_equal(buf.ind, ind), key
    for size in (cap, 1234, 1):
        sb = oreplay.OracleUniform(cap, "Hopper-v4")
        sb.size = size
        sb.sample(256, u=torch.tensor(g[f"uni/{size}/u"]))
        assert np.array_equal(sb.ind, g[f"uni/{size}/ind"])


def test_replay_big_priority_array():
    g = gu.load("replay_cases")
    big = (np.maximum(np.abs(np.random.default_rng(11).standard_normal(1_000_000) * 3), 1.0) ** 0.4).astype(np.float32)
    _close(gu.checksum(big), g["big/priority_ck"])
    assert np.array_equal(oreplay.sample_indices_numpy(big, g["big/u"]), g["big/ind"])


def test_replay_append_wrap_and_scatter():
    g = gu.load("replay_cases")
    buf = oreplay.OracleLAP(3000, "Hopper-v4")
    for i in range(3500):
        buf.append([g["wrap/obs"][i], g["wrap/act"][i], float(g["wrap/rew"][i]), g["wrap/next_obs"][i], float(g["wrap/notdone"][i])])
        if i == 1700:
            buf.sample(64, u=torch.tensor(g["wrap/mid_u"]))
            assert np.array_equal(buf.ind, g["wrap/mid_ind"])
            buf.update_priority(torch.tensor(g["wrap/mid_prio"]))
    assert [buf.ptr, buf.size, len(buf)] == list(g["wrap/ptr_size_len"])
    _close(buf.priority.numpy(), g["wrap/priority"])
    _close(float(buf.max_priority), float(g["wrap/max_priority"]))
    _close(gu.checksum(buf.state), g["wrap/state_ck"])
    _close(gu.checksum(buf.action), g["wrap/action_ck"])
    # duplicate scatter / running max / reset
    buf = oreplay.OracleLAP(3000, "Hopper-v4")
    buf.priority[:] = 1.0
    buf.size = 3000
    buf.ind = g["dup/ind"]
    buf.update_priority(torch.tensor(g["dup/prio"]))
    _close(buf.priority.numpy(), g["dup/priority_after"])
    _close(float(buf.max_priority), float(g["dup/max_priority"]))
    buf.priority[100] = 1.0
    buf.reset_max_priority()
    _close(float(buf.max_priority), float(g["dup/max_after_reset"]))
