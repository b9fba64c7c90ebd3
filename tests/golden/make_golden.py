"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures, so parity is pinned
against outputs of the reference itself: this script imports ``rl`` from
``/root/reference`` under ``oracle.gym_stub`` and records, for a few update
steps of each algorithm, everything the hot path produces:

* the uniforms drawn by ``sample`` and the Gaussian noise drawn by the agent
  (captured by wrapping ``torch.rand`` / ``torch.randn_like`` / ``Tensor.normal_``),
* sampled indices, the info dict, priorities written back, ``max_priority``,
* fp64 checksums (sum, |sum|, head) of every gradient and every weight tensor.

Inputs are NOT stored: buffer rows come from ``oracle.replay.synthetic_rows(seed)``
and initial weights from the constructors under ``torch.manual_seed(seed)``; the
checksums of the initial weights are stored so a regeneration mismatch is caught.
``/root/reference`` does not exist on the GPU box; only the .npz files travel.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gym_stub  # noqa: E402
from oracle.replay import synthetic_rows  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (algo, env, prioritized buffer, agent kwargs, steps)
    "td7_halfcheetah_lap": ("TD7", "HalfCheetah-v4", True, dict(use_lap=True, target_update_rate=4), 9),
    "td7_hopper_uniform": ("TD7", "Hopper-v4", False, dict(use_lap=False, target_update_rate=3), 5),
    "td3_hopper_uniform": ("TD3", "Hopper-v4", False, dict(use_lap=False), 5),
    "td3_ant_lap": ("TD3", "Ant-v4", True, dict(use_lap=True), 4),
    "sac_walker2d": ("SAC", "Walker2d-v4", False, dict(), 5),
    "td7_humanoid_lap": ("TD7", "Humanoid-v4", True, dict(use_lap=True, target_update_rate=250), 2),
}
BUFFER_ROWS = 4096
BATCH = 256
SEED = 1234


def checksum(t: torch.Tensor) -> np.ndarray:
    x = t.detach().double().reshape(-1)
    head = torch.zeros(8, dtype=torch.float64)
    n = min(8, x.numel())
    head[:n] = x[:n]
    return np.concatenate([[x.sum().item(), x.abs().sum().item(), float(x.numel())], head.numpy()])


class Capture:
    """Record every draw of the three RNG entry points the hot path uses."""

    def __init__(self):
        self.rand, self.randn = [], []
        self._rand, self._randn_like, self._normal_ = torch.rand, torch.randn_like, torch.Tensor.normal_

    def __enter__(self):
        cap = self

        def rand(*a, **k):
            v = cap._rand(*a, **k)
            cap.rand.append(v.clone())
            return v

        def randn_like(*a, **k):
            v = cap._randn_like(*a, **k)
            cap.randn.append(v.clone())
            return v

        def normal_(self_t, *a, **k):
            v = cap._normal_(self_t, *a, **k)
            cap.randn.append(v.clone())
            return v

        torch.rand, torch.randn_like, torch.Tensor.normal_ = rand, randn_like, normal_
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn_like, torch.Tensor.normal_ = self._rand, self._randn_like, self._normal_


def module_sets(agent, algo):
    names = ["policy", "q1", "q2", "target_q1", "target_q2"]
    if algo == "TD7":
        names += ["encoder", "fixed_encoder", "fixed_encoder_target"]
    return {n: getattr(agent, n) for n in names}


def run_case(name, algo, env_id, prioritized, kwargs, steps):
    rl = gym_stub.import_reference()
    from rl.agent import SAC, TD3, TD7
    from rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    s_dim, a_dim, _ = gym_stub.ENV_SPECS[env_id]
    torch.manual_seed(SEED)
    agent = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo](env_id, **kwargs)
    buf = (LAPReplayMemory if prioritized else SimpleReplayMemory)(BUFFER_ROWS, env_id)
    rows = synthetic_rows(BUFFER_ROWS - 96, s_dim, a_dim, seed=SEED)  # size < capacity on purpose
    n = rows["state"].shape[0]
    buf.state[:n], buf.action[:n], buf.reward[:n] = rows["state"], rows["action"], rows["reward"]
    buf.next_state[:n], buf.float_done[:n] = rows["next_state"], rows["done"]
    buf.size, buf.ptr = n, n
    if prioritized:
        buf.priority[:n] = 1.0

    out = {"meta_steps": np.array(steps), "meta_rows": np.array(n)}
    for mod_name, mod in module_sets(agent, algo).items():
        for k, v in mod.state_dict().items():
            out[f"init/{mod_name}/{k}"] = checksum(v)

    for t in range(steps):
        with Capture() as cap:
            batch = buf.sample(BATCH)
            info = agent.train_ops(batch, replay_buffer=buf)
        assert len(cap.rand) == 1
        out[f"s{t}/u"] = cap.rand[0].numpy()
        for j, z in enumerate(cap.randn):
            out[f"s{t}/noise{j}"] = z.numpy()
        out[f"s{t}/n_noise"] = np.array(len(cap.randn))
        out[f"s{t}/ind"] = np.asarray(buf.ind, dtype=np.int64)
        out[f"s{t}/batch_state_ck"] = checksum(batch["state"])
        out[f"s{t}/batch_next_state_ck"] = checksum(batch["next_state"])
        out[f"s{t}/batch_action"] = batch["action"].numpy()
        out[f"s{t}/batch_reward"] = batch["reward"].numpy()
        out[f"s{t}/batch_done"] = batch["done"].numpy()
        for k, v in info.items():
            out[f"s{t}/info/{k}"] = np.array(np.nan if v is None else v, dtype=np.float64)
        if prioritized:
            out[f"s{t}/priority_at_ind"] = buf.priority[buf.ind].numpy().copy()
            out[f"s{t}/max_priority"] = np.array(float(buf.max_priority))
            out[f"s{t}/priority_ck"] = checksum(buf.priority[: buf.size])
        if algo == "TD7":
            out[f"s{t}/value_range"] = np.array(
                [float(agent.value_max), float(agent.value_min), float(agent.value_target_max), float(agent.value_target_min)]
            )
        for mod_name, mod in module_sets(agent, algo).items():
            for k, v in mod.state_dict().items():
                out[f"s{t}/w/{mod_name}/{k}"] = checksum(v)
        if algo == "SAC":
            out[f"s{t}/tmp"] = np.array(float(agent.tmp.detach()))
    # acting path (deterministic) on a fixed observation
    obs = np.linspace(-1.0, 1.0, s_dim).astype(np.float32)
    out["act/obs"] = obs
    out["act/action"] = np.asarray(agent.sample(obs, True), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok", {k: out[f"s{steps-1}/info/{k}"].item() for k in info})


def replay_cases():
    """LAP sampler edge cases recorded from the reference."""
    gym_stub.import_reference()
    from rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    out = {}
    g = np.random.default_rng(7)
    env_id = "Hopper-v4"
    s_dim, a_dim, _ = gym_stub.ENV_SPECS[env_id]
    cap = 3000
    buf = LAPReplayMemory(cap, env_id)
    # ring wrap through append(): 3500 appends into a 3000-row ring
    tr = []
    for i in range(3500):
        t = [g.standard_normal(s_dim).astype(np.float32), g.uniform(-1, 1, a_dim).astype(np.float32),
             float(np.float32(g.standard_normal())), g.standard_normal(s_dim).astype(np.float32), float(i % 50 != 0)]
        tr.append(t)
        buf.append([x.copy() if hasattr(x, "copy") else x for x in t])
        if i == 1700:  # priorities move mid-stream so appended rows pick up a non-unit max_priority
            u = torch.rand(64)
            _orig = torch.rand
            torch.rand = lambda *a, **k: u
            b = buf.sample(64)
            torch.rand = _orig
            pr = torch.tensor(g.uniform(1.0, 9.0, 64).astype(np.float32))
            buf.update_priority(pr)
            out["wrap/mid_u"], out["wrap/mid_prio"] = u.numpy(), pr.numpy()
            out["wrap/mid_ind"] = np.asarray(buf.ind, dtype=np.int64)
    out["wrap/obs"] = np.stack([t[0] for t in tr])
    out["wrap/act"] = np.stack([t[1] for t in tr])
    out["wrap/rew"] = np.array([t[2] for t in tr], dtype=np.float32)
    out["wrap/next_obs"] = np.stack([t[3] for t in tr])
    out["wrap/notdone"] = np.array([t[4] for t in tr], dtype=np.float32)
    out["wrap/ptr_size_len"] = np.array([buf.ptr, buf.size, len(buf)])
    out["wrap/priority"] = buf.priority.numpy().copy()
    out["wrap/max_priority"] = np.array(float(buf.max_priority))
    out["wrap/state_ck"] = checksum(torch.tensor(buf.state))
    out["wrap/action_ck"] = checksum(torch.tensor(buf.action))

    # sampling against assorted priority arrays with injected uniforms incl. edge values
    edge_u = np.array([0.0, 1.0 - 2.0**-24, 0.5, 2.0**-24, 0.25, 0.75], dtype=np.float32)
    pri_sets = {
        "ones": np.ones(cap, np.float32),
        "rand": g.uniform(1.0, 30.0, cap).astype(np.float32),
        "zeros_mixed": np.where(g.random(cap) < 0.3, 0.0, g.uniform(1.0, 5.0, cap)).astype(np.float32),
        "pow": (np.maximum(np.abs(g.standard_normal(cap) * 4), 1.0) ** 0.4).astype(np.float32),
    }
    for nm, pri in pri_sets.items():
        for size in (cap, 1777, 1):
            buf.priority[:] = torch.tensor(pri)
            buf.size = size
            u = np.concatenate([edge_u, g.random(250, dtype=np.float32)]).astype(np.float32)
            ut = torch.tensor(u)
            _orig = torch.rand
            torch.rand = lambda *a, **k: ut
            b = buf.sample(256)
            torch.rand = _orig
            key = f"samp/{nm}/{size}"
            out[key + "/priority"] = pri
            out[key + "/u"] = u
            out[key + "/ind"] = np.asarray(buf.ind, dtype=np.int64)
            out[key + "/state_ck"] = checksum(b["state"])
    # duplicate-index scatter: last batch position wins; running max; reset
    buf.priority[:] = 1.0
    buf.size = cap
    buf.max_priority = 1
    buf.ind = np.array([5, 9, 5, 5, 100, 9, 2999, 0], dtype=np.int64)
    pr = torch.tensor([2.0, 3.0, 4.0, 1.5, 7.0, 2.5, 6.0, 1.0])
    buf.update_priority(pr)
    out["dup/ind"], out["dup/prio"] = buf.ind, pr.numpy()
    out["dup/priority_after"] = buf.priority.numpy().copy()
    out["dup/max_priority"] = np.array(float(buf.max_priority))
    buf.priority[100] = 1.0
    buf.reset_max_priority()
    out["dup/max_after_reset"] = np.array(float(buf.max_priority))

    # uniform buffer
    sb = SimpleReplayMemory(cap, env_id)
    for size in (cap, 1234, 1):
        sb.size = size
        u = np.concatenate([edge_u, g.random(250, dtype=np.float32)]).astype(np.float32)
        ut = torch.tensor(u)
        _orig = torch.rand
        torch.rand = lambda *a, **k: ut
        sb.sample(256)
        torch.rand = _orig
        out[f"uni/{size}/u"] = u
        out[f"uni/{size}/ind"] = np.asarray(sb.ind, dtype=np.int64)
    # a 1M-row array (the size BASELINE.json quotes): only u, indices and a generator seed stored
    big = (np.maximum(np.abs(np.random.default_rng(11).standard_normal(1_000_000) * 3), 1.0) ** 0.4).astype(np.float32)
    lb = LAPReplayMemory(1_000_000, env_id)
    lb.priority[:] = torch.tensor(big)
    lb.size = 1_000_000
    u = np.concatenate([edge_u, g.random(250, dtype=np.float32)]).astype(np.float32)
    ut = torch.tensor(u)
    _orig = torch.rand
    torch.rand = lambda *a, **k: ut
    lb.sample(256)
    torch.rand = _orig
    out["big/u"], out["big/ind"] = u, np.asarray(lb.ind, dtype=np.int64)
    out["big/priority_ck"] = checksum(torch.tensor(big))
    np.savez_compressed(os.path.join(OUT, "replay_cases.npz"), **out)
    print("replay_cases ok")


if __name__ == "__main__":
    torch.set_num_threads(1)
    replay_cases()
    for nm, (algo, env_id, pri, kw, steps) in CASES.items():
        run_case(nm, algo, env_id, pri, kw, steps)
