"""GPU-resident ring replay: ``LAPReplayMemory`` / ``SimpleReplayMemory`` drop-ins.

Same constructor, methods and side channels as ``rl/replay_memory/lap.py`` / ``simple.py``
(``append``, ``sample``, ``update_priority``, ``reset_max_priority``, ``len()`` == ptr, ``.size``,
``.max_priority``, ``.ind``), backed by the CUDA kernels of ``csrc/replay.cuh`` through the C ABI.
``sample`` returns CUDA tensors; the uniforms are drawn on the device (Philox) unless ``u`` is
given.  Extra keyword arguments (all optional): ``state_dim`` / ``action_dim`` instead of
``env_id``, ``n_seeds`` (stack independent seeds), ``share_rows`` (one transition store for all
stacked seeds, per-seed priorities).
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from ... import _lib
from ...envs import resolve
from .base import BaseReplayMemory


_stream = _lib.current_stream_ptr


class _DeviceRing(BaseReplayMemory):
    prioritized = True

    def __init__(self, replay_buffer_size: int, env_id: str | None = None, *, state_dim=None, action_dim=None,
                 n_seeds: int = 1, share_rows: bool = False, device="cuda", zero_copy: bool = True, **kwargs) -> None:
        super().__init__(replay_buffer_size, **kwargs)
        if not torch.cuda.is_available():
            raise _lib.B200RLError("sac_td3_td7_b200 replay memory needs a CUDA device (no CPU fallback)")
        self._lib = _lib.lib()
        self.device = torch.device(device)
        self.state_dim, self.action_dim, self.action_bias, self.action_scale = resolve(env_id, state_dim, action_dim)
        self.n_seeds, self.share_rows = int(n_seeds), bool(share_rows)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200rl_replay_create(int(replay_buffer_size), self.state_dim, self.action_dim, self.n_seeds,
                                                      int(self.prioritized), int(self.share_rows), C.byref(h)))
        self._h = h
        self.row_stride = self._lib.b200rl_replay_row_stride(self._h)
        offs = (C.c_int * 5)()
        _lib.check(self._lib.b200rl_replay_row_layout(self._h, offs))
        self.row_offsets = dict(zip(("state", "action", "reward", "done", "next_state"), offs))  # float offsets inside a packed row
        self.ptr = 0
        self.size = 0
        self._last_batch = 0
        # lap.py:35 runs on the device: the append kernel divides / subtracts in IEEE fp32 (numpy's result for fp32 actions)
        A = self.action_dim
        self._scale32 = np.ascontiguousarray(np.broadcast_to(np.asarray(self.action_scale, np.float32), (A,)))
        self._bias32 = np.ascontiguousarray(np.broadcast_to(np.asarray(self.action_bias, np.float32), (A,)))
        _lib.check(self._lib.b200rl_replay_set_action_norm(self._h, self._scale32.ctypes.data, self._bias32.ctypes.data))
        self._rew1, self._nd1 = np.zeros(1, np.float32), np.zeros(1, np.float32)
        self._rew1_p, self._nd1_p = self._rew1.ctypes.data, self._nd1.ctypes.data
        # zero-copy hand-off: once an agent has trained from this buffer (train_ops(batch, replay_buffer=self)),
        # sample() writes straight into that agent's fixed input buffers and returns tensors that ALIAS them --
        # the next sample() overwrites them.  zero_copy=False restores private tensors per call.
        self.zero_copy = bool(zero_copy)
        self._sink = None  # weakref to the agent whose input buffers sample() fills

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.b200rl_replay_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # ---- producer side (rl/rollout.py:45) -------------------------------------------------
    def pack(self, state, action, reward, next_state, float_done) -> np.ndarray:
        """Packed fp32 rows [n, row_stride] from already-normalised columns."""
        state = np.asarray(state, np.float32).reshape(-1, self.state_dim)
        n, S, A = state.shape[0], self.state_dim, self.action_dim
        o = self.row_offsets
        rows = np.zeros((n, self.row_stride), np.float32)
        rows[:, o["state"]:o["state"] + S] = state
        rows[:, o["action"]:o["action"] + A] = np.asarray(action, np.float32).reshape(n, A)
        rows[:, o["reward"]] = np.asarray(reward, np.float32).reshape(n)
        rows[:, o["done"]] = np.asarray(float_done, np.float32).reshape(n)
        rows[:, o["next_state"]:o["next_state"] + S] = np.asarray(next_state, np.float32).reshape(n, S)
        return rows

    def unpack(self, rows: np.ndarray) -> dict:
        """The five columns of packed rows [n, row_stride] (inverse of pack)."""
        S, A, o = self.state_dim, self.action_dim, self.row_offsets
        return dict(state=rows[:, o["state"]:o["state"] + S], action=rows[:, o["action"]:o["action"] + A], reward=rows[:, o["reward"]:o["reward"] + 1],
                    done=rows[:, o["done"]:o["done"] + 1], next_state=rows[:, o["next_state"]:o["next_state"] + S])

    def append(self, transition: list[Any], seed: int = 0) -> None:
        """lap.py:31-43 -- one transition [obs, action(env units), reward, next_obs, notdone].

        One C call: the library packs the row into a pinned staging slot and ONE kernel copies it into the ring,
        normalises the action (``a / scale - bias``, lap.py:35), writes the priority and repairs the sum tree.  fp32
        actions (what ``agent.sample`` and gym's ``action_space.sample`` produce) are normalised on the device in IEEE fp32
        -- the value numpy computes; any other dtype is normalised here the way numpy would and stored."""
        assert len(transition) == 5
        obs, action, reward, next_obs, float_done = transition
        action = np.asarray(action)
        raw = action.dtype == np.float32
        if not raw:
            action = np.asarray(action / self.action_scale - self.action_bias, np.float32)  # (sic) lap.py:35
        obs = np.ascontiguousarray(obs, np.float32)
        action = np.ascontiguousarray(action)
        next_obs = np.ascontiguousarray(next_obs, np.float32)
        assert obs.size == self.state_dim and next_obs.size == self.state_dim and action.size == self.action_dim
        self._rew1[0], self._nd1[0] = reward, float_done
        _lib.check(self._lib.b200rl_replay_append_transitions(self._h, seed, 1, obs.ctypes.data, action.ctypes.data, self._rew1_p, next_obs.ctypes.data,
                                                              self._nd1_p, int(raw), _stream()))
        if seed == 0:
            self.ptr = (self.ptr + 1) % self.replay_buffer_size
            self.size = min(self.size + 1, self.replay_buffer_size)

    def append_population(self, obs, action, reward, next_obs, float_done) -> None:
        """One environment step of every stacked seed (rl/rollout.py:45 once per seed) in ONE launch: obs / next_obs
        [n_seeds, S], action [n_seeds, A] in environment units, reward / float_done [n_seeds].  With ``share_rows`` the
        store keeps seed 0's transition and every seed's priorities / cursor advance."""
        G = self.n_seeds
        action = np.asarray(action)
        raw = action.dtype == np.float32
        if not raw:
            action = np.asarray(action / self.action_scale - self.action_bias, np.float32)
        obs = np.ascontiguousarray(obs, np.float32).reshape(G, self.state_dim)
        action = np.ascontiguousarray(action).reshape(G, self.action_dim)
        next_obs = np.ascontiguousarray(next_obs, np.float32).reshape(G, self.state_dim)
        reward = np.ascontiguousarray(reward, np.float32).reshape(G)
        float_done = np.ascontiguousarray(float_done, np.float32).reshape(G)
        _lib.check(self._lib.b200rl_replay_append_transitions(self._h, 0, G, obs.ctypes.data, action.ctypes.data, reward.ctypes.data, next_obs.ctypes.data,
                                                              float_done.ctypes.data, int(raw), _stream()))
        self.ptr = (self.ptr + 1) % self.replay_buffer_size
        self.size = min(self.size + 1, self.replay_buffer_size)

    def append_rows(self, rows, seed: int = 0) -> None:
        """Bulk append of packed rows (numpy / CPU tensor -> H2D, CUDA tensor -> D2D).

        Up to 64 rows whose memory the GPU can read directly -- CUDA tensors, or numpy views of PINNED host memory --
        are appended by one kernel that reads the source when it executes, in stream order (like cudaMemcpyAsync from
        pinned memory: keep such a source unchanged until then, e.g. rotate a few pinned slots).  Pageable rows are
        staged by the driver at call time and may be reused immediately."""
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            rows = rows.contiguous().float()
            _lib.check(self._lib.b200rl_replay_append_device(self._h, seed, rows.data_ptr(), rows.shape[0], _stream()))
            n = rows.shape[0]
        else:
            rows = np.ascontiguousarray(rows.numpy() if isinstance(rows, torch.Tensor) else rows, np.float32)
            _lib.check(self._lib.b200rl_replay_append(self._h, seed, rows.ctypes.data, rows.shape[0], _stream()))
            n = rows.shape[0]
        if seed == 0:
            self.ptr = (self.ptr + n) % self.replay_buffer_size
            self.size = min(self.size + n, self.replay_buffer_size)

    def fill_synthetic(self, n: int, seed: int = 0, chunk: int = 65536) -> None:
        """BASELINE.md section 2 synthetic contents (the generator bench.py and the oracle share)."""
        g = np.random.default_rng(seed)
        S, A = self.state_dim, self.action_dim
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            st = g.standard_normal((m, S), dtype=np.float32)
            ac = g.random((m, A), dtype=np.float32) * 2.0 - 1.0
            rw = g.standard_normal((m, 1), dtype=np.float32)
            ns = g.standard_normal((m, S), dtype=np.float32)
            dn = (g.random((m, 1), dtype=np.float32) < 0.99).astype(np.float32)
            rows = self.pack(st, ac, rw, ns, dn)
            if self.n_seeds > 1 and not self.share_rows:
                rows = torch.from_numpy(rows).to(self.device)  # one H2D, then a device-to-device copy into every seed's private store
            for s in range(1 if self.share_rows else self.n_seeds):
                self.append_rows(rows, s)
            if self.share_rows:
                for s in range(1, self.n_seeds):  # cursor + priorities of the other stacked seeds
                    self.append_rows(rows, s)

    # ---- consumer side (rl/runner/run.py:93) ----------------------------------------------
    def sample(self, batch_size: int, use_torch: bool = True, u: torch.Tensor | None = None, out: dict | None = None):
        """lap.py:45-64.  Returns fp32 CUDA tensors [B, dim] ([n_seeds, B, dim] when stacked)."""
        B, G, S, A = int(batch_size), self.n_seeds, self.state_dim, self.action_dim
        sink = self._sink() if self._sink is not None else None
        if sink is not None and out is None and use_torch and sink._h and B == sink.batch_size and sink._inbox is not None:
            u_ptr = None
            if u is not None:
                u = torch.as_tensor(u, dtype=torch.float32).to(self.device).contiguous()
                assert u.numel() == G * B
                u_ptr = u.data_ptr()
            _lib.check(self._lib.b200rl_replay_sample_to_agent(self._h, sink._h, B, u_ptr, _stream()))
            self._last_batch = B
            return sink._inbox
        if out is None:
            shape = (lambda d: (B, d)) if G == 1 else (lambda d: (G, B, d))
            out = dict(state=torch.empty(shape(S), device=self.device), action=torch.empty(shape(A), device=self.device),
                       reward=torch.empty(shape(1), device=self.device), next_state=torch.empty(shape(S), device=self.device),
                       done=torch.empty(shape(1), device=self.device))
        u_ptr = None
        if u is not None:
            u = torch.as_tensor(u, dtype=torch.float32).to(self.device).contiguous()
            assert u.numel() == G * B
            u_ptr = u.data_ptr()
        _lib.check(self._lib.b200rl_replay_sample(self._h, B, u_ptr, out["state"].data_ptr(), out["action"].data_ptr(), out["reward"].data_ptr(),
                                                  out["next_state"].data_ptr(), out["done"].data_ptr(), _stream()))
        self._last_batch = B
        if not use_torch:
            return {k: v.cpu().numpy() for k, v in out.items()}
        return out

    @property
    def ind(self) -> np.ndarray:
        """Indices of the most recent sample (lap.py:61); copied from the device on access."""
        idx = np.empty((self.n_seeds, self._last_batch), np.int64)
        _lib.check(self._lib.b200rl_replay_last_indices(self._h, idx.ctypes.data, _stream()))
        return idx[0] if self.n_seeds == 1 else idx

    def __len__(self) -> int:
        return self.ptr  # lap.py:75-76 (sic: ptr, not size)


class LAPReplayMemory(_DeviceRing):
    """Drop-in for ``rl.replay_memory.LAPReplayMemory`` (lap.py:12)."""

    prioritized = True

    def update_priority(self, priority: torch.Tensor) -> None:
        """lap.py:66-69 -- scatter at the last sampled indices (duplicates: last wins) + running max."""
        p = torch.as_tensor(priority, dtype=torch.float32).to(self.device).contiguous().view(-1)
        B = p.numel() // self.n_seeds
        _lib.check(self._lib.b200rl_replay_update_priority(self._h, p.data_ptr(), B, _stream()))

    def reset_max_priority(self) -> None:
        """lap.py:71-73"""
        _lib.check(self._lib.b200rl_replay_reset_max_priority(self._h, _stream()))

    @property
    def max_priority(self) -> float:
        m = C.c_float()
        _lib.check(self._lib.b200rl_replay_get_state(self._h, 0, None, None, C.byref(m), _stream()))
        return float(m.value)

    @property
    def priority(self) -> torch.Tensor:
        """Host copy of priority[:capacity] of seed 0 (inspection / tests)."""
        out = np.empty(self.replay_buffer_size, np.float32)
        _lib.check(self._lib.b200rl_replay_read_priority(self._h, 0, 0, self.replay_buffer_size, out.ctypes.data, _stream()))
        return torch.from_numpy(out)


class SimpleReplayMemory(_DeviceRing):
    """Drop-in for ``rl.replay_memory.SimpleReplayMemory`` (simple.py:12): uniform sampling."""

    prioritized = False
