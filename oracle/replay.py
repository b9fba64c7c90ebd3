"""Oracle: ring replay with LAP / uniform sampling.  TEST INFRASTRUCTURE ONLY.

Restates ``rl/replay_memory/lap.py`` and ``rl/replay_memory/simple.py``:

* storage: five numpy float64 arrays + a float32 priority vector (lap.py:23-29)
* append:  ``a/scale - bias``, write row ``ptr``, ``priority[ptr]=max_priority``,
           advance ring (lap.py:31-43)
* sample:  inclusive prefix sum of ``priority[:size]`` (float32 in/out, float64
           accumulate -- torch CPU semantics), ``v = u * total`` in float32,
           lower-bound search, gather rows, cast to float32 (lap.py:45-64)
* update_priority: scatter (duplicates: last batch position wins) and running
           max (lap.py:66-69);  reset_max_priority (lap.py:71-73)
* ``len()`` is ``ptr`` (lap.py:75-76)

``sample`` takes an optional ``u`` (float32 uniforms in [0,1)) so the CUDA path
and the oracle can be fed identical draws; with ``u=None`` it draws
``torch.rand(B)`` from the global CPU generator exactly where the reference does.
"""
from __future__ import annotations

import numpy as np
import torch

from . import envs


def prefix_sums_f32(priority: np.ndarray) -> np.ndarray:
    """float32 inclusive prefix sums with float64 accumulation (numpy restatement).

    torch.cumsum on a CPU float32 tensor accumulates in double and rounds every
    prefix to float32; this is the same thing in numpy.
    """
    return np.cumsum(priority.astype(np.float64)).astype(np.float32)


def lower_bound_indices(prefix: np.ndarray, value: np.ndarray) -> np.ndarray:
    """First i with prefix[i] >= value (torch.searchsorted(right=False))."""
    return np.searchsorted(prefix, value, side="left").astype(np.int64)


def sample_indices_numpy(priority: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Pure-numpy restatement of lap.py:47-54 for given uniforms."""
    prefix = prefix_sums_f32(priority)
    value = (u.astype(np.float32) * prefix[-1]).astype(np.float32)
    return lower_bound_indices(prefix, value)


class _Ring:
    def __init__(self, replay_buffer_size: int, env_id: str | None = None, *, state_dim=None, action_dim=None):
        self.replay_buffer_size = int(replay_buffer_size)
        if env_id is not None:
            state_dim, action_dim = envs.dims(env_id)
            self.action_bias, self.action_scale = envs.action_bias_scale(env_id)
        else:
            self.action_bias = np.zeros((action_dim,), np.float32)
            self.action_scale = np.ones((action_dim,), np.float32)
        n = self.replay_buffer_size
        self.state = np.zeros((n, state_dim))
        self.action = np.zeros((n, action_dim))
        self.next_state = np.zeros((n, state_dim))
        self.reward = np.zeros((n, 1))
        self.float_done = np.zeros((n, 1))
        self.ptr = 0
        self.size = 0
        self.ind = None

    def _write(self, transition) -> int:
        assert len(transition) == 5
        obs, action, reward, next_obs, float_done = transition
        action = action / self.action_scale - self.action_bias
        p = self.ptr
        self.state[p] = obs
        self.action[p] = action
        self.reward[p] = reward
        self.next_state[p] = next_obs
        self.float_done[p] = float_done
        self.ptr = (p + 1) % self.replay_buffer_size
        self.size = min(self.size + 1, self.replay_buffer_size)
        return p

    def _gather(self, ind: np.ndarray) -> dict:
        f = lambda arr: torch.Tensor(arr[ind])  # float64 -> float32 cast, as the reference
        return dict(
            state=f(self.state),
            action=f(self.action),
            reward=f(self.reward),
            next_state=f(self.next_state),
            done=f(self.float_done),
        )

    def fill_synthetic(self, n: int, seed: int = 0) -> None:
        """BASELINE.md section 2 synthetic contents (same generator as the CUDA bench)."""
        rows = synthetic_rows(n, self.state.shape[1], self.action.shape[1], seed)
        self.state[:n] = rows["state"]
        self.action[:n] = rows["action"]
        self.reward[:n] = rows["reward"]
        self.next_state[:n] = rows["next_state"]
        self.float_done[:n] = rows["done"]
        self.size = min(n, self.replay_buffer_size)
        self.ptr = n % self.replay_buffer_size

    def __len__(self) -> int:
        return self.ptr


def synthetic_rows(n: int, state_dim: int, action_dim: int, seed: int = 0) -> dict:
    """Synthetic transitions: s,s'~N(0,1); a~U(-1,1); r~N(0,1); notdone~Bernoulli(0.99).

    float32 values (so float64 oracle storage and float32 GPU storage hold the
    same numbers).  Generated blockwise so a 1M x 376 buffer does not need a
    second full-size temporary.
    """
    g = np.random.default_rng(seed)
    out = {
        "state": np.empty((n, state_dim), np.float32),
        "action": np.empty((n, action_dim), np.float32),
        "reward": np.empty((n, 1), np.float32),
        "next_state": np.empty((n, state_dim), np.float32),
        "done": np.empty((n, 1), np.float32),
    }
    step = 65536
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        m = hi - lo
        out["state"][lo:hi] = g.standard_normal((m, state_dim), dtype=np.float32)
        out["action"][lo:hi] = g.random((m, action_dim), dtype=np.float32) * 2.0 - 1.0
        out["reward"][lo:hi] = g.standard_normal((m, 1), dtype=np.float32)
        out["next_state"][lo:hi] = g.standard_normal((m, state_dim), dtype=np.float32)
        out["done"][lo:hi] = (g.random((m, 1), dtype=np.float32) < 0.99).astype(np.float32)
    return out


class OracleLAP(_Ring):
    """Restatement of ``LAPReplayMemory`` (lap.py:12-76)."""

    prioritized = True

    def __init__(self, replay_buffer_size: int, env_id: str | None = None, **kw):
        super().__init__(replay_buffer_size, env_id, **kw)
        self.priority = torch.zeros(self.replay_buffer_size)
        self.max_priority = 1

    def append(self, transition) -> None:
        p = self._write(transition)
        self.priority[p] = self.max_priority

    def fill_synthetic(self, n: int, seed: int = 0) -> None:
        super().fill_synthetic(n, seed)
        self.priority[: self.size] = 1.0

    def sample(self, batch_size: int, use_torch: bool = True, u: torch.Tensor | None = None) -> dict:
        prefix = torch.cumsum(self.priority[: self.size], 0)
        if u is None:
            u = torch.rand(batch_size)
        value = u * prefix[-1]
        ind = torch.searchsorted(prefix, value).cpu().data.numpy()
        self.ind = ind
        return self._gather(ind)

    def update_priority(self, priority: torch.Tensor) -> None:
        self.priority[self.ind] = priority
        self.max_priority = max(float(priority.max()), self.max_priority)

    def reset_max_priority(self) -> None:
        self.max_priority = float(self.priority[: self.size].max())


class OracleUniform(_Ring):
    """Restatement of ``SimpleReplayMemory`` (simple.py:12-65): LAP with all-ones priorities."""

    prioritized = False

    def append(self, transition) -> None:
        self._write(transition)

    def sample(self, batch_size: int, use_torch: bool = True, u: torch.Tensor | None = None) -> dict:
        prefix = torch.cumsum(torch.ones(self.size), 0)
        if u is None:
            u = torch.rand(batch_size)
        value = u * prefix[-1]
        ind = torch.searchsorted(prefix, value).cpu().data.numpy()
        self.ind = ind
        return self._gather(ind)
