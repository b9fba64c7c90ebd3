"""Oracle: SALE and MLP networks as pure functions of a parameter dict.  TEST INFRASTRUCTURE ONLY.

Restates ``rl/nn/sale.py`` and ``rl/nn/mlp.py``.  Parameters live in plain
``dict[str, Tensor]`` whose keys equal the reference modules' ``state_dict()``
keys, so weights move between the reference, the oracle and the CUDA path by
name.  ``init_*`` draw from the global torch CPU generator in the same order as
the reference constructors, so ``torch.manual_seed(s)`` gives identical weights.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

Params = dict  # name -> Tensor


# ----------------------------------------------------------------------------- init
def _linear(params: Params, name: str, fan_in: int, fan_out: int, xavier: bool = False) -> None:
    # nn.Linear's default init consumes the generator exactly like the reference
    # (kaiming_uniform(a=sqrt 5) on weight, U(+-1/sqrt(fan_in)) on bias).
    lin = nn.Linear(fan_in, fan_out)
    if xavier:  # rl/nn/mlp.py:20-30: xavier_normal_ weights, zero bias, applied after construction
        nn.init.xavier_normal_(lin.weight.data)
        nn.init.zeros_(lin.bias.data)
    params[name + ".weight"] = lin.weight.data.clone()
    params[name + ".bias"] = lin.bias.data.clone()


def init_sale_encoder(state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256) -> Params:
    """sale.py:19-39."""
    p: Params = {}
    _linear(p, "zs1", state_dim, hdim)
    _linear(p, "zs2", hdim, hdim)
    _linear(p, "zs3", hdim, zs_dim)
    _linear(p, "zsa1", zs_dim + action_dim, hdim)
    _linear(p, "zsa2", hdim, hdim)
    _linear(p, "zsa3", hdim, zs_dim)
    return p


def init_sale_actor(state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256) -> Params:
    """sale.py:61-75."""
    p: Params = {}
    _linear(p, "l0", state_dim, hdim)
    _linear(p, "l1", zs_dim + hdim, hdim)
    _linear(p, "l2", hdim, hdim)
    _linear(p, "l3", hdim, action_dim)
    return p


def init_sale_critic(state_dim: int, action_dim: int, zs_dim: int = 256, hdim: int = 256) -> Params:
    """sale.py:89-104."""
    p: Params = {}
    _linear(p, "q01", state_dim + action_dim, hdim)
    _linear(p, "q1", 2 * zs_dim + hdim, hdim)
    _linear(p, "q2", hdim, hdim)
    _linear(p, "q3", hdim, 1)
    return p


def init_mlp(input_dim: int, output_dim: int, hidden: int = 256) -> Params:
    """mlp.py:10-35 (Linear-ReLU-Linear-ReLU-Linear; keys follow nn.Sequential: mlp.0/2/4)."""
    p: Params = {}
    _linear(p, "mlp.0", input_dim, hidden, xavier=True)
    _linear(p, "mlp.2", hidden, hidden, xavier=True)
    _linear(p, "mlp.4", hidden, output_dim, xavier=True)
    return p


# ----------------------------------------------------------------------------- forward
def avg_l1_norm(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """sale.py:11-13."""
    return x / x.abs().mean(-1, keepdim=True).clamp(min=eps)


def _lin(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def encode_state(p: Params, state: torch.Tensor) -> torch.Tensor:
    """sale.py:41-46."""
    h = F.elu(_lin(p, "zs1", state))
    h = F.elu(_lin(p, "zs2", h))
    return avg_l1_norm(_lin(p, "zs3", h))


def encode_state_action(p: Params, zs: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """sale.py:48-55."""
    h = F.elu(_lin(p, "zsa1", torch.cat([zs, action], 1)))
    h = F.elu(_lin(p, "zsa2", h))
    return _lin(p, "zsa3", h)


def sale_actor(p: Params, obs: torch.Tensor, zs: torch.Tensor) -> torch.Tensor:
    """sale.py:77-83."""
    a = avg_l1_norm(_lin(p, "l0", obs))
    a = torch.cat([a, zs], 1)
    a = F.relu(_lin(p, "l1", a))
    a = F.relu(_lin(p, "l2", a))
    return torch.tanh(_lin(p, "l3", a))


def sale_critic(p: Params, obs, action, zsa, zs) -> torch.Tensor:
    """sale.py:106-121."""
    sa = torch.cat([obs, action], 1)
    emb = torch.cat([zsa, zs], 1)
    q = avg_l1_norm(_lin(p, "q01", sa))
    q = torch.cat([q, emb], 1)
    q = F.elu(_lin(p, "q1", q))
    q = F.elu(_lin(p, "q2", q))
    return _lin(p, "q3", q)


def mlp(p: Params, x: torch.Tensor) -> torch.Tensor:
    """mlp.py:24-35."""
    h = F.relu(_lin(p, "mlp.0", x))
    h = F.relu(_lin(p, "mlp.2", h))
    return _lin(p, "mlp.4", h)


def mlp_critic(p: Params, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """mlp.py:99-102."""
    return mlp(p, torch.cat([state, action], -1))


def clone_params(p: Params, requires_grad: bool = False) -> Params:
    return {k: v.detach().clone().requires_grad_(requires_grad) for k, v in p.items()}
