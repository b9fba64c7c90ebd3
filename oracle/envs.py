"""Shapes and action bounds of the MuJoCo tasks named in BASELINE.json (oracle side).

The reference obtains these by instantiating a gymnasium env
(``rl/utils/miscellaneous.py:50-66``); gymnasium/MuJoCo are absent here, so the
oracle carries the table.  bias = (low+high)/2, scale = (high-low)/2.
"""
from __future__ import annotations

import numpy as np

ENV_SPECS = {
    "Hopper-v4": (11, 3, 1.0),
    "HalfCheetah-v4": (17, 6, 1.0),
    "Walker2d-v4": (17, 6, 1.0),
    "Ant-v4": (27, 8, 1.0),
    "Humanoid-v4": (376, 17, 0.4),
}


def dims(env_id: str) -> tuple[int, int]:
    s, a, _ = ENV_SPECS[env_id]
    return s, a


def action_bias_scale(env_id: str) -> tuple[np.ndarray, np.ndarray]:
    _, a, bound = ENV_SPECS[env_id]
    low = np.full((a,), -bound, dtype=np.float32)
    high = np.full((a,), bound, dtype=np.float32)
    return (low + high) / 2.0, (high - low) / 2.0
