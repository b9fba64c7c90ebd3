"""TEST INFRASTRUCTURE ONLY -- fake ``gymnasium`` so the reference imports without MuJoCo.

Used by ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) to import the *unmodified* reference and record
golden vectors. The reference only needs gymnasium for env shapes and action
bounds on this path (``rl/utils/miscellaneous.py:50-66``) plus a few base-class
names at import time (``rl/utils/make_env.py:5``, ``rl/runner/run.py:7-9``).

Nothing in the product package imports this module.
"""
from __future__ import annotations

import sys
import types

import numpy as np

# obs dim, act dim, action bound  (MuJoCo v4 tasks named in BASELINE.json)
ENV_SPECS = {
    "Hopper-v4": (11, 3, 1.0),
    "HalfCheetah-v4": (17, 6, 1.0),
    "Walker2d-v4": (17, 6, 1.0),
    "Ant-v4": (27, 8, 1.0),
    "Humanoid-v4": (376, 17, 0.4),
}


class _Box:
    def __init__(self, dim: int, bound: float):
        self.shape = (dim,)
        self.low = np.full((dim,), -bound, dtype=np.float32)
        self.high = np.full((dim,), bound, dtype=np.float32)

    def seed(self, *_a, **_k):
        return None

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


class _Env:
    def __init__(self, env_id: str):
        s, a, bound = ENV_SPECS[env_id]
        self.observation_space = _Box(s, np.inf)
        self.action_space = _Box(a, bound)
        self.spec = types.SimpleNamespace(id=env_id)

    def reset(self, **_k):
        return np.zeros(self.observation_space.shape, dtype=np.float32), {}


def install() -> None:
    """Register the stub modules in ``sys.modules`` (idempotent)."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_b200_stub", False):
        return
    gym = types.ModuleType("gymnasium")
    gym._b200_stub = True
    gym.registry = {k: None for k in ENV_SPECS}
    gym.make = lambda env_id, **kw: _Env(getattr(env_id, "id", env_id))
    gym.Env = type("Env", (), {})
    gym.Space = type("Space", (), {})
    gym.ActionWrapper = type("ActionWrapper", (), {"__init__": lambda self, env: None})

    utils = types.ModuleType("gymnasium.utils")
    utils.RecordConstructorArgs = type("RecordConstructorArgs", (), {})
    gym.utils = utils

    wrappers = types.ModuleType("gymnasium.wrappers")
    wrappers.TimeLimit = type("TimeLimit", (), {})
    wrappers.FlattenObservation = type("FlattenObservation", (), {})
    res = types.ModuleType("gymnasium.wrappers.record_episode_statistics")
    res.RecordEpisodeStatistics = type("RecordEpisodeStatistics", (), {})
    rv = types.ModuleType("gymnasium.wrappers.record_video")
    rv.RecordVideo = type("RecordVideo", (), {})
    wrappers.record_episode_statistics = res
    wrappers.record_video = rv
    gym.wrappers = wrappers

    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.utils"] = utils
    sys.modules["gymnasium.wrappers"] = wrappers
    sys.modules["gymnasium.wrappers.record_episode_statistics"] = res
    sys.modules["gymnasium.wrappers.record_video"] = rv
    # optional third-party modules the reference's non-hot-path files import
    for name in ("click_repl", "moviepy", "moviepy.editor", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                m = types.ModuleType(name)
                m.register_repl = lambda *a, **k: None
                sys.modules[name] = m


def import_reference(path: str = "/root/reference"):
    """Import the reference package ``rl`` from ``path`` under the stub."""
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    import rl  # noqa: F401
    import rl.agent  # noqa: F401

    return sys.modules["rl"]
