"""Shared helpers to replay the golden cases recorded by tests/golden/make_golden.py."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror tests/golden/make_golden.py
CASES = {
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


def load(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def checksum(t) -> np.ndarray:
    x = torch.as_tensor(t).detach().double().reshape(-1).cpu()
    head = torch.zeros(8, dtype=torch.float64)
    n = min(8, x.numel())
    head[:n] = x[:n]
    return np.concatenate([[x.sum().item(), x.abs().sum().item(), float(x.numel())], head.numpy()])


def make_oracle(name: str):
    """Oracle agent + buffer initialised exactly as the reference was for this case."""
    from oracle import agents, envs, replay

    algo, env_id, prioritized, kwargs, steps = CASES[name]
    torch.manual_seed(SEED)
    agent = {"TD7": agents.OracleTD7, "TD3": agents.OracleTD3, "SAC": agents.OracleSAC}[algo](env_id, **kwargs)
    buf = (replay.OracleLAP if prioritized else replay.OracleUniform)(BUFFER_ROWS, env_id)
    buf.fill_synthetic(BUFFER_ROWS - 96, seed=SEED)
    return agent, buf, algo, steps


def step_noise(g, t: int, algo: str):
    n = int(g[f"s{t}/n_noise"])
    zs = [torch.tensor(g[f"s{t}/noise{j}"]) for j in range(n)]
    if algo == "SAC":
        return (zs[0], zs[1])
    return zs[0]
