"""Observation / action shapes of the MuJoCo tasks (``rl/utils/miscellaneous.py:50-66``).

The reference instantiates a gymnasium env to read them; gymnasium / MuJoCo are optional here:
when importable they are used (same code path as the reference), otherwise the table below
serves the five tasks of BASELINE.json.  Explicit ``state_dim`` / ``action_dim`` always work.
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


def _gym_env(env_id: str):
    try:
        import gymnasium as gym  # noqa: PLC0415
    except Exception:  # noqa: BLE001
        return None
    if getattr(gym, "_b200_stub", False) or env_id not in getattr(gym, "registry", {}):
        return None
    try:
        return gym.make(env_id)
    except Exception:  # noqa: BLE001
        return None


def get_state_action_dims(env_id: str) -> tuple[int, int]:
    env = _gym_env(env_id)
    if env is not None:
        return env.observation_space.shape[0], env.action_space.shape[0]
    if env_id not in ENV_SPECS:
        raise KeyError(f"unknown env_id {env_id!r}: install gymnasium or pass state_dim / action_dim")
    s, a, _ = ENV_SPECS[env_id]
    return s, a


def get_action_bias_scale(env_id: str) -> tuple[np.ndarray, np.ndarray]:
    env = _gym_env(env_id)
    if env is not None:
        low, high = env.action_space.low, env.action_space.high
    else:
        _, a, bound = ENV_SPECS[env_id]
        low = np.full((a,), -bound, dtype=np.float32)
        high = np.full((a,), bound, dtype=np.float32)
    return (low + high) / 2.0, (high - low) / 2.0


def resolve(env_id, state_dim, action_dim):
    """(S, A, action_bias, action_scale) from an env id or explicit dims."""
    if env_id is not None:
        s, a = get_state_action_dims(env_id)
        bias, scale = get_action_bias_scale(env_id)
        return s, a, bias, scale
    if state_dim is None or action_dim is None:
        raise ValueError("pass env_id or both state_dim and action_dim")
    return int(state_dim), int(action_dim), np.zeros((action_dim,), np.float32), np.ones((action_dim,), np.float32)
