"""``rl.nn`` modules: SALE (TD7) and MLP (TD3/SAC) networks.

Same constructors, parameter names and initialisation as ``rl/nn/sale.py`` / ``rl/nn/mlp.py`` (they
hold real ``nn.Linear`` parameters, so ``torch.manual_seed`` gives the reference's weights and
``state_dict`` is interchangeable).  The forward methods run the fused CUDA kernels (inference
only); training happens inside the agents' update engines, which import these weights.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as Fn
from .abc import ACTOR, CRITIC, ENCODER


def AvgL1Norm(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """sale.py:11-13"""
    return Fn.avg_l1_norm(x, eps)


def _act_name(fn) -> str:
    for name, ref in (("elu", F.elu), ("relu", F.relu), ("tanh", torch.tanh)):
        if fn is ref:
            return name
    raise NotImplementedError("only F.elu / F.relu / torch.tanh activations have fused kernels")


class SALEEncoder(ENCODER):
    """sale.py:16-55"""

    def __init__(self, state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256, activ: Callable = F.elu):
        super().__init__()
        self.activ = activ
        self.zs1 = nn.Linear(state_dim, hdim)
        self.zs2 = nn.Linear(hdim, hdim)
        self.zs3 = nn.Linear(hdim, zs_dim)
        self.zsa1 = nn.Linear(zs_dim + action_dim, hdim)
        self.zsa2 = nn.Linear(hdim, hdim)
        self.zsa3 = nn.Linear(hdim, zs_dim)

    def encode_state(self, state: torch.Tensor) -> torch.Tensor:
        a = _act_name(self.activ)
        zs = Fn.linear(state, self.zs1.weight, self.zs1.bias, a)
        zs = Fn.linear(zs, self.zs2.weight, self.zs2.bias, a)
        return Fn.avg_l1_norm(Fn.linear(zs, self.zs3.weight, self.zs3.bias))

    def encode_state_action(self, zs: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        a = _act_name(self.activ)
        zsa = Fn.linear([zs, action], self.zsa1.weight, self.zsa1.bias, a)
        zsa = Fn.linear(zsa, self.zsa2.weight, self.zsa2.bias, a)
        return Fn.linear(zsa, self.zsa3.weight, self.zsa3.bias)


class SALEActor(ACTOR):
    """sale.py:58-83"""

    def __init__(self, state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256, activ: Callable = F.relu):
        super().__init__()
        self.activ = activ
        self.l0 = nn.Linear(state_dim, hdim)
        self.l1 = nn.Linear(zs_dim + hdim, hdim)
        self.l2 = nn.Linear(hdim, hdim)
        self.l3 = nn.Linear(hdim, action_dim)

    def inference_mean(self, obs: torch.Tensor, zs: torch.Tensor) -> torch.Tensor:
        act = _act_name(self.activ)
        a = Fn.avg_l1_norm(Fn.linear(obs, self.l0.weight, self.l0.bias))
        a = Fn.linear([a, zs], self.l1.weight, self.l1.bias, act)
        a = Fn.linear(a, self.l2.weight, self.l2.bias, act)
        return Fn.linear(a, self.l3.weight, self.l3.bias, "tanh")


class SALECritic(CRITIC):
    """sale.py:86-121"""

    def __init__(self, state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256, activ: Callable = F.elu):
        super().__init__()
        self.activ = activ
        self.q01 = nn.Linear(state_dim + action_dim, hdim)
        self.q1 = nn.Linear(2 * zs_dim + hdim, hdim)
        self.q2 = nn.Linear(hdim, hdim)
        self.q3 = nn.Linear(hdim, 1)

    def estimate_q_value(self, obs, action, zsa, zs) -> torch.Tensor:
        act = _act_name(self.activ)
        q = Fn.avg_l1_norm(Fn.linear([obs, action], self.q01.weight, self.q01.bias))
        q = Fn.linear([q, zsa, zs], self.q1.weight, self.q1.bias, act)
        q = Fn.linear(q, self.q2.weight, self.q2.bias, act)
        return Fn.linear(q, self.q3.weight, self.q3.bias)


def make_mlp(input_dim: int, output_dim: int, hidden_sizes: list[int] = [256, 256], action_fn: str | None = "ReLU",
             init_weight: None | str = None, init_bias: None | str = None) -> nn.Module:
    """mlp.py:10-35 (a plain nn.Sequential; used for parameters / state_dict)."""
    init_w = nn.init.xavier_normal_ if init_weight is None else getattr(nn.init, init_weight)
    init_b = nn.init.zeros_ if init_bias is None else getattr(nn.init, init_bias)
    fn = getattr(nn, action_fn)() if isinstance(action_fn, str) else action_fn
    mods: list[nn.Module] = []
    for i, o in zip([input_dim] + hidden_sizes, hidden_sizes + [output_dim]):
        lin = nn.Linear(i, o)
        init_w(lin.weight.data)
        init_b(lin.bias.data)
        mods.append(lin)
        if fn is not None:
            mods.append(fn)
    mods.pop(-1)
    return nn.Sequential(*mods)


def _mlp_forward(seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    lins = [m for m in seq if isinstance(m, nn.Linear)]
    for i, lin in enumerate(lins):
        x = Fn.linear(x, lin.weight, lin.bias, "relu" if i + 1 < len(lins) else None)
    return x


class MLPActor(ACTOR):
    """mlp.py:38-72"""

    def __init__(self, state_dim: int, action_dim: int, hidden_sizes: list[int] | int = 256, **mlp_kwargs) -> None:
        super().__init__()
        hidden_sizes = [hidden_sizes] * 2 if isinstance(hidden_sizes, int) else hidden_sizes
        self.mlp = make_mlp(input_dim=state_dim, output_dim=action_dim, hidden_sizes=hidden_sizes, **mlp_kwargs)

    def inference_mean(self, state) -> torch.Tensor:
        return _mlp_forward(self.mlp, state)

    def inference_mean_logvar(self, state):
        mean, logvar = _mlp_forward(self.mlp, state).chunk(2, -1)
        return mean, logvar

    def inference_policy(self, state) -> torch.Tensor:
        return _mlp_forward(self.mlp, state)


class MLPCritic(CRITIC):
    """mlp.py:75-104"""

    def __init__(self, state_dim: int, action_dim: int, hidden_sizes: list[int] | int = 256, **mlp_kwargs) -> None:
        super().__init__()
        hidden_sizes = [hidden_sizes] * 2 if isinstance(hidden_sizes, int) else hidden_sizes
        self.mlp = make_mlp(input_dim=state_dim + action_dim, output_dim=1, hidden_sizes=hidden_sizes, **mlp_kwargs)

    def estimate_q_value(self, state, action) -> torch.Tensor:
        lins = [m for m in self.mlp if isinstance(m, nn.Linear)]
        x = Fn.linear([state, action], lins[0].weight, lins[0].bias, "relu")
        for i, lin in enumerate(lins[1:], 1):
            x = Fn.linear(x, lin.weight, lin.bias, "relu" if i + 1 < len(lins) else None)
        return x

    def estimate_state_value(self, state) -> torch.Tensor:
        return _mlp_forward(self.mlp, state)
