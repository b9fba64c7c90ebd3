"""TD7 drop-in (``rl/agent/td7.py:31-335``): same constructor and methods, update on the GPU."""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import Callable

import numpy as np
import torch

from ... import _lib
from ...envs import resolve
from ..nn import SALEActor, SALECritic, SALEEncoder
from ..replay_memory import LAPReplayMemory
from .abc import Agent, LazyScalar, _stream

_ENC = ["zs1", "zs2", "zs3", "zsa1", "zsa2", "zsa3"]
_ACT = ["l0", "l1", "l2", "l3"]
_CRI = ["q01", "q1", "q2", "q3"]


class TD7(Agent):
    """Agent for TD7."""

    NETS = [("encoder", _ENC), ("policy", _ACT), ("q1", _CRI), ("q2", _CRI), ("target_q1", _CRI), ("target_q2", _CRI),
            ("fixed_encoder", _ENC), ("fixed_encoder_target", _ENC)]
    INFO_KEYS = ["train/encoder", "train/q_fn", "train/policy"]
    TRAINED = {"encoder", "policy", "q1", "q2"}

    def __init__(self, env_id: str | None = None, discount_factor: float = 0.99, policy_lr: float = 3e-4, critic_lr: float = 3e-4,
                 target_update_rate: int = 250, exploration_noise: float = 0.1, target_policy_noise: float = 0.2, noise_clip: float = 0.5,
                 policy_freq: int = 2, use_lap: bool = False, make_nn: Callable | None = None, *, state_dim=None, action_dim=None,
                 batch_size: int = 256, n_seeds: int = 1, zs_dim: int = 256, hdim: int = 256, keep_grads: bool = False, sync_info: bool = False,
                 device="cuda", **make_nn_kwargs) -> None:
        S, A, self.action_bias, self.action_scale = resolve(env_id, state_dim, action_dim)
        self.state_dim, self.action_dim = S, A
        if make_nn is None:
            policy, q1, q2, encoder = self.make_nn(S, A, zs_dim, hdim)
        else:  # td7.py:56-61: custom factory; only the SALE architecture has fused kernels
            kw = {k: v for k, v in make_nn_kwargs.items() if k not in ("state_dim", "action_dim")}
            policy, q1, q2, encoder = make_nn(state_dim=S, action_dim=A, **kw)
            if not (isinstance(policy, SALEActor) and isinstance(q1, SALECritic) and isinstance(encoder, SALEEncoder)):
                raise NotImplementedError("the fused update engine supports the SALE networks of rl.nn only")
            hdim, zs_dim = encoder.zs2.out_features, encoder.zs3.out_features
        self.zs_dim, self.hdim = zs_dim, hdim
        # td7.py:62-66 (target_policy aliases policy: td7.py:103)
        self._mods = dict(encoder=encoder, policy=policy, q1=q1, q2=q2, target_q1=deepcopy(q1), target_q2=deepcopy(q2),
                          fixed_encoder=deepcopy(encoder), fixed_encoder_target=deepcopy(encoder))
        self.discount_factor = discount_factor
        self.policy_lr, self.critic_lr = policy_lr, critic_lr
        self.target_update_rate = target_update_rate
        self.target_policy_noise = target_policy_noise
        self.exploration_noise = exploration_noise
        self.noise_clip = noise_clip
        self.policy_freq = policy_freq
        self.use_lap = use_lap
        self.batch_size, self.n_seeds = int(batch_size), int(n_seeds)
        self.keep_grads, self.sync_info = bool(keep_grads), bool(sync_info)
        self._n_calls = 0
        self._act_out_dim = A
        self._engine_init(device)

    @staticmethod
    def make_nn(state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256):
        """td7.py:89-98 (same construction order, so a shared torch seed gives the same weights)."""
        encoder = SALEEncoder(state_dim, action_dim, zs_dim, hdim)
        policy = SALEActor(state_dim, action_dim, zs_dim, hdim)
        q1 = SALECritic(state_dim, action_dim, zs_dim, hdim)
        q2 = SALECritic(state_dim, action_dim, zs_dim, hdim)
        return policy, q1, q2, encoder

    def _fresh_modules(self) -> dict:
        policy, q1, q2, encoder = self.make_nn(self.state_dim, self.action_dim, self.zs_dim, self.hdim)
        return dict(encoder=encoder, policy=policy, q1=q1, q2=q2, target_q1=q1, target_q2=q2, fixed_encoder=encoder, fixed_encoder_target=encoder)

    def _create_engine(self):
        cfg = _lib.TD7Config(state_dim=self.state_dim, action_dim=self.action_dim, hdim=self.hdim, zs_dim=self.zs_dim, batch=self.batch_size,
                             groups=self.n_seeds, discount=self.discount_factor, policy_lr=self.policy_lr, critic_lr=self.critic_lr,
                             target_update_rate=self.target_update_rate, policy_freq=self.policy_freq, target_policy_noise=self.target_policy_noise,
                             noise_clip=self.noise_clip, use_lap=int(self.use_lap), keep_grads=int(self.keep_grads))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200rl_td7_create(C.byref(cfg), C.byref(h)))
        self._h = h

    # networks as attributes, like the reference (synced from the device on access)
    policy = property(lambda self: self._module("policy"))
    target_policy = property(lambda self: self._module("policy"))
    q1 = property(lambda self: self._module("q1"))
    q2 = property(lambda self: self._module("q2"))
    target_q1 = property(lambda self: self._module("target_q1"))
    target_q2 = property(lambda self: self._module("target_q2"))
    encoder = property(lambda self: self._module("encoder"))
    fixed_encoder = property(lambda self: self._module("fixed_encoder"))
    fixed_encoder_target = property(lambda self: self._module("fixed_encoder_target"))

    @property
    def n_runs(self) -> int:
        return self._n_calls

    def _range(self, seed: int = 0):
        """[value_max, value_min, value_target_max, value_target_min] of one stacked seed; seed = -1: [n_seeds, 4]."""
        out = np.empty(4 if seed >= 0 else (self.n_seeds, 4), np.float32)
        _lib.check(self._lib.b200rl_td7_value_range(self._h, seed, out.ctypes.data, _stream()))
        return out

    value_max = property(lambda self: float(self._range()[0]))
    value_min = property(lambda self: float(self._range()[1]))
    value_target_max = property(lambda self: float(self._range()[2]))
    value_target_min = property(lambda self: float(self._range()[3]))

    def _extra_state(self):
        return {"__range__": self._range(-1)}

    def _restore_extra_state(self, extra):
        if "__range__" in extra:
            r = np.ascontiguousarray(extra["__range__"], np.float32)
            _lib.check(self._lib.b200rl_td7_set_value_range(self._h, -1 if r.ndim == 2 else 0, r.ctypes.data, _stream()))

    def _actor_step(self, call_index: int) -> bool:
        return call_index % self.policy_freq == 0  # td7.py:295,319: n_runs is incremented first

    def _info_dict(self, step: int, actor: bool) -> dict:
        vals = [LazyScalar(self, step, c) for c in range(3)]
        if self.sync_info:
            vals = [float(v) for v in vals]
        return {"train/encoder": vals[0], "train/q_fn": vals[1], "train/policy": vals[2] if actor else None}

    def train_ops(self, batch, replay_buffer: LAPReplayMemory | None = None, *args, noise=None, **kwargs) -> dict:
        """td7.py:287-332.  `noise` optionally injects the N(0,1) draws of the target smoothing."""
        if self.use_lap:
            assert isinstance(replay_buffer, LAPReplayMemory)  # td7.py:310
        self._stage(batch)
        self._update(replay_buffer if self.use_lap else None, noise, sink=replay_buffer)
        return self._info_dict(self._n_calls, self._actor_step(self._n_calls))

    def sample(self, state, deterministic: bool = False, **kwargs) -> np.ndarray:
        """td7.py:141-162"""
        action = self._policy_out(state)
        if not deterministic:
            action = action + (torch.randn(action.shape) * self.exploration_noise).numpy()
        action = np.clip(action, -1.0, 1.0)
        return action * self.action_scale + self.action_bias

    def __repr__(self) -> str:
        return "TD7"
