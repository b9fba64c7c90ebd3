"""TD3 drop-in (``rl/agent/td3.py:30-245``)."""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import Callable

import numpy as np
import torch

from ... import _lib
from ...envs import resolve
from ..nn import MLPActor, MLPCritic
from ..replay_memory import LAPReplayMemory
from .abc import Agent, LazyScalar

_MLP = ["mlp.0", "mlp.2", "mlp.4"]


class _MlpAgent(Agent):
    NETS = [("policy", _MLP), ("q1", _MLP), ("q2", _MLP), ("target_q1", _MLP), ("target_q2", _MLP)]
    TRAINED = {"policy", "q1", "q2"}
    ALGO = 0

    def _mlp_cfg(self):
        return _lib.MlpConfig(algo=self.ALGO, state_dim=self.state_dim, action_dim=self.action_dim, hdim=self.hdim, batch=self.batch_size,
                              groups=self.n_seeds, discount=self.discount_factor, policy_lr=self.policy_lr, critic_lr=self.critic_lr, tau=self.tau,
                              policy_freq=getattr(self, "policy_freq", 1), target_policy_noise=getattr(self, "target_policy_noise", 0.0),
                              noise_clip=getattr(self, "noise_clip", 0.0), min_log_std=getattr(self, "min_log_std", -20.0),
                              max_log_std=getattr(self, "max_log_std", 2.0), tmp=float(getattr(self, "_tmp_init", -1.0)),
                              use_lap=int(self.use_lap), keep_grads=int(self.keep_grads))

    def _create_engine(self):
        cfg = self._mlp_cfg()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200rl_mlp_create(C.byref(cfg), C.byref(h)))
        self._h = h

    policy = property(lambda self: self._module("policy"))
    q1 = property(lambda self: self._module("q1"))
    q2 = property(lambda self: self._module("q2"))
    target_q1 = property(lambda self: self._module("target_q1"))
    target_q2 = property(lambda self: self._module("target_q2"))

    @property
    def n_runs(self) -> int:
        return self._n_calls

    @staticmethod
    def _hidden(policy) -> int:
        return policy.mlp[0].out_features


class TD3(_MlpAgent):
    """Agent for TD3."""

    INFO_KEYS = ["train/q_fn", "train/policy", "norm/policy"]
    ALGO = 0

    def __init__(self, env_id: str | None = None, discount_factor: float = 0.99, policy_lr: float = 3e-4, critic_lr: float = 3e-4,
                 exploration_noise: float = 0.1, target_policy_noise: float = 0.2, noise_clip: float = 0.5, policy_freq: int = 2, tau: float = 0.005,
                 use_lap: bool = False, make_nn: Callable | None = None, *, state_dim=None, action_dim=None, batch_size: int = 256, n_seeds: int = 1,
                 keep_grads: bool = False, sync_info: bool = False, device="cuda", **make_nn_kwargs) -> None:
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
        self.discount_factor = discount_factor
        self.policy_lr, self.critic_lr = policy_lr, critic_lr
        self.target_policy_noise = target_policy_noise
        self.exploration_noise = exploration_noise
        self.noise_clip = noise_clip
        self.policy_freq = policy_freq
        self.use_lap = use_lap
        self.tau = tau
        self.batch_size, self.n_seeds = int(batch_size), int(n_seeds)
        self.keep_grads, self.sync_info = bool(keep_grads), bool(sync_info)
        self._n_calls = 0
        self._act_out_dim = A
        self._engine_init(device)

    target_policy = property(lambda self: self._module("policy"))  # aliased (td3.py:86)

    def _fresh_modules(self) -> dict:
        policy, q1, q2 = self.make_nn(self.state_dim, self.action_dim)
        return dict(policy=policy, q1=q1, q2=q2, target_q1=q1, target_q2=q2)

    def make_nn(self, state_dim: int, action_dim: int):
        """td3.py:76-81"""
        return MLPActor(state_dim, action_dim), MLPCritic(state_dim, action_dim), MLPCritic(state_dim, action_dim)

    def _actor_step(self, call_index: int) -> bool:
        return (call_index - 1) % self.policy_freq == 0  # td3.py:231,241: n_runs is incremented last

    def _info_dict(self, step: int, actor: bool) -> dict:
        vals = [LazyScalar(self, step, c) for c in range(3)]
        if self.sync_info:
            vals = [float(v) for v in vals]
        return {"train/q_fn": vals[0], "train/policy": vals[1] if actor else None, "norm/policy": vals[2] if actor else None}

    def train_ops(self, batch, replay_buffer=None, *args, noise=None, **kwargs) -> dict:
        """td3.py:206-242"""
        if self.use_lap:
            assert isinstance(replay_buffer, LAPReplayMemory)
        self._stage(batch)
        self._update(replay_buffer if self.use_lap else None, noise, sink=replay_buffer)
        return self._info_dict(self._n_calls, self._actor_step(self._n_calls))

    def sample(self, state, deterministic: bool = False, **kwargs) -> np.ndarray:
        """td3.py:114-141"""
        action = self._policy_out(state)
        if not deterministic:
            action = action + (torch.randn(action.shape) * self.exploration_noise).numpy()
        action = np.clip(action, -1.0, 1.0)
        return action * self.action_scale + self.action_bias

    def __repr__(self) -> str:
        return "TD3"
