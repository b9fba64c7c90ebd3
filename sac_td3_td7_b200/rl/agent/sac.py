"""SAC drop-in (``rl/agent/sac.py:24-299``), automatic or fixed temperature."""
from __future__ import annotations

from copy import deepcopy
from typing import Callable

import numpy as np
import torch

from ... import _lib
from ...envs import resolve
from ..nn import MLPActor, MLPCritic
from .abc import LazyScalar, _stream
from .td3 import _MlpAgent


class SAC(_MlpAgent):
    """Agent for Soft Actor Critic."""

    INFO_KEYS = ["train/q_fn", "tmp", "norm/tmp", "train/policy", "train/tmp", "entropy"]
    ALGO = 1

    def __init__(self, env_id: str | None = None, discount_factor: float = 0.99, policy_lr: float = 3e-4, critic_lr: float = 3e-4,
                 min_log_std: float = -20.0, max_log_std: float = 2.0, tau: float = 0.005, tmp: float = -1.0, use_lap: bool = False,
                 max_grad_norm: float = float("inf"), make_nn: Callable | None = None, *, state_dim=None, action_dim=None, batch_size: int = 256,
                 n_seeds: int = 1, keep_grads: bool = False, sync_info: bool = False, device="cuda", **make_nn_kwargs) -> None:
        S, A, self.action_bias, self.action_scale = resolve(env_id, state_dim, action_dim)
        self.state_dim, self.action_dim = S, A
        if make_nn is None:
            policy, q1, q2 = self.make_nn(S, A)
        else:
            kw = {k: v for k, v in make_nn_kwargs.items() if k not in ("state_dim", "action_dim")}
            policy, q1, q2 = make_nn(state_dim=S, action_dim=A, **kw)
            if not (isinstance(policy, MLPActor) and isinstance(q1, MLPCritic)):
                raise NotImplementedError("the fused update engine supports the 2-hidden-layer MLP networks of rl.nn only")
        self.hdim = self._hidden(policy)
        self._mods = dict(policy=policy, q1=q1, q2=q2, target_q1=deepcopy(q1), target_q2=deepcopy(q2))
        self.auto_tmp_mode = tmp < 0.0
        self._tmp_init = tmp
        self.target_entropy = -A  # sac.py:67
        self.discount_factor = discount_factor
        self.policy_lr, self.critic_lr = policy_lr, critic_lr
        self.min_log_std, self.max_log_std = min_log_std, max_log_std
        self.tau = tau
        self.use_lap = False  # SAC+LAP is dead code in the reference; a LAP buffer just samples with unit priorities
        self.max_grad_norm = max_grad_norm  # stored, never used (sac.py:76)
        self.batch_size, self.n_seeds = int(batch_size), int(n_seeds)
        self.keep_grads, self.sync_info = bool(keep_grads), bool(sync_info)
        self._n_calls = 0
        self._act_out_dim = 2 * A
        self._engine_init(device)

    @staticmethod
    def make_nn(state_dim: int, action_dim: int):
        """sac.py:80-86"""
        return MLPActor(state_dim, action_dim * 2), MLPCritic(state_dim, action_dim), MLPCritic(state_dim, action_dim)

    def _fresh_modules(self) -> dict:
        policy, q1, q2 = self.make_nn(self.state_dim, self.action_dim)
        return dict(policy=policy, q1=q1, q2=q2, target_q1=q1, target_q2=q2)

    @property
    def tmp(self):
        """log-temperature parameter in auto mode (sac.py:55-60), else the fixed temperature."""
        if not self.auto_tmp_mode:
            return self._tmp_init
        out = np.empty(1, np.float32)
        _lib.check(self._lib.b200rl_agent_read(self._h, 5, 0, 0, out.ctypes.data, _stream()))
        return torch.tensor(out)

    def _extra_state(self):
        if not self.auto_tmp_mode:
            return {}
        out = np.empty((3, self.n_seeds), np.float32)  # log-temperature, exp_avg, exp_avg_sq of every stacked seed
        for w in range(3):
            _lib.check(self._lib.b200rl_agent_read(self._h, 5, w, -1, out[w].ctypes.data, _stream()))
        return {"__tmp__": out}

    def _restore_extra_state(self, extra):
        if "__tmp__" in extra:
            for w in range(3):
                v = np.ascontiguousarray(extra["__tmp__"][w], np.float32).reshape(-1)
                _lib.check(self._lib.b200rl_agent_write(self._h, 5, w, -1 if v.size == self.n_seeds and self.n_seeds > 1 else 0, v.ctypes.data, _stream()))

    def _info_dict(self, step: int, actor: bool) -> dict:
        vals = [LazyScalar(self, step, c) for c in range(6)]
        if self.sync_info:
            vals = [float(v) for v in vals]
        info = {"train/q_fn": vals[0]}
        if self.auto_tmp_mode:
            info["tmp"], info["norm/tmp"] = vals[1], vals[2]
        info["train/policy"] = vals[3]
        if self.auto_tmp_mode:
            info["train/tmp"] = vals[4]
        info["entropy"] = vals[5]
        return info

    def train_ops(self, batch, replay_buffer=None, *args, noise=None, **kwargs) -> dict:
        """sac.py:251-295.  `noise` = stacked (critic draw, actor draw) N(0,1) tensors [2, B, A]."""
        self._stage(batch)
        if noise is not None and isinstance(noise, (tuple, list)):
            noise = torch.stack([torch.as_tensor(noise[0]), torch.as_tensor(noise[1])])
        self._update(None, noise, sink=replay_buffer)
        return self._info_dict(self._n_calls, True)

    @torch.no_grad()
    def sample(self, state, deterministic: bool = False, **kwargs) -> np.ndarray:
        """sac.py:131-159"""
        out = self._policy_out(state)  # [2A], or [n_seeds, 2A] for a population
        A = self.action_dim
        mean, log_std = out[..., :A], np.clip(out[..., A:], self.min_log_std, self.max_log_std)
        if deterministic:
            action = np.tanh(mean)
        else:
            action = np.tanh(mean + torch.empty(mean.shape).normal_().numpy() * np.exp(log_std))
        return action * self.action_scale + self.action_bias

    def __repr__(self) -> str:
        return "SAC"
