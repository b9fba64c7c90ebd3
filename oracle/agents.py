"""Oracle: TD7 / TD3 / SAC update steps on CPU.  TEST INFRASTRUCTURE ONLY.

Restates ``train_ops`` of ``rl/agent/td7.py:287-332``, ``rl/agent/td3.py:206-242``
and ``rl/agent/sac.py:251-295`` over the functional nets of ``oracle.nets``.
Dense math, autograd and Adam are PyTorch's (third-party for the reference too:
torch==2.0.1 pinned in its requirements.txt, torch 2.11 available here).

Differences from the reference that do not change any number:
* Gaussian noise can be injected (``noise=...``) so the CUDA path and the oracle
  see the same draws; when omitted the oracle calls ``torch.randn_like`` /
  ``torch.empty(...).normal_()`` at the same points, in the same order, as the
  reference, so a shared ``torch.manual_seed`` reproduces the reference stream.
* gradients that the reference computes and never uses are not
  computed.
* every step leaves its intermediate results in ``self.last`` for the parity tests.

Reference quirks that DO change numbers are kept:
target_policy aliases policy (B.1), value clip is [0,0] before the first target
sync (B.2), step-counter phases (B.3), TD3's Polyak of the policy onto itself
(B.1), SAC logging ``train/policy`` = policy + temperature loss (B.8).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import envs, nets

EPS = 1e-6  # rl/utils/annotation.py:12


def _leafs(p: nets.Params) -> nets.Params:
    return nets.clone_params(p, requires_grad=True)


def _frozen(p: nets.Params) -> nets.Params:
    return nets.clone_params(p, requires_grad=False)


def _copy_into(dst: nets.Params, src: nets.Params) -> None:
    with torch.no_grad():
        for k in dst:
            dst[k].copy_(src[k])


def _grads(p: nets.Params) -> dict:
    return {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in p.items()}


def grad_norm_sum(p: nets.Params) -> float:
    """rl/nn/utils.py:13-19 -- sum over tensors of the L2 norm of each gradient."""
    total = 0.0
    for v in p.values():
        if v.grad is not None:
            total += torch.norm(v.grad.data, p=2)
    return total.item()


def adam_step_restated(p, g, m, v, step: int, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8):
    """Single-tensor Adam exactly as torch/optim/adam.py `_single_tensor_adam` orders it.

    In place on float32 tensors ``p, m, v``; ``step`` is the 1-based step count.
    Used to pin the CUDA Adam kernel in isolation (identical input gradients).
    """
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


def lap_huber_sum(td: torch.Tensor, min_priority: float = 1.0) -> torch.Tensor:
    """td7.py:165-173 -- sum over the two critics, mean over the batch."""
    return torch.where(td < min_priority, 0.5 * td.pow(2), min_priority * td).sum(1).mean()


def lap_huber_mean(td: torch.Tensor, min_priority: float = 1.0) -> torch.Tensor:
    """td3.py:137-142 -- plain mean."""
    return torch.where(td < min_priority, 0.5 * td.pow(2), min_priority * td).mean()


class _Base:
    def _dims(self, env_id, state_dim, action_dim):
        if env_id is not None:
            state_dim, action_dim = envs.dims(env_id)
            self.action_bias, self.action_scale = envs.action_bias_scale(env_id)
        else:
            self.action_bias = np.zeros((action_dim,), np.float32)
            self.action_scale = np.ones((action_dim,), np.float32)
        self.state_dim, self.action_dim = state_dim, action_dim
        return state_dim, action_dim


# =============================================================================== TD7
class OracleTD7(_Base):
    """Restatement of ``TD7`` (td7.py:31-335)."""

    def __init__(self, env_id=None, discount_factor=0.99, policy_lr=3e-4, critic_lr=3e-4,
                 target_update_rate=250, exploration_noise=0.1, target_policy_noise=0.2,
                 noise_clip=0.5, policy_freq=2, use_lap=False, *, state_dim=None, action_dim=None,
                 zs_dim=256, hdim=256, **_ignored):
        s, a = self._dims(env_id, state_dim, action_dim)
        # construction order of td7.py:90-98: encoder, policy, q1, q2
        self.encoder = _leafs(nets.init_sale_encoder(s, a, zs_dim, hdim))
        self.policy = _leafs(nets.init_sale_actor(s, a, zs_dim, hdim))
        self.q1 = _leafs(nets.init_sale_critic(s, a, zs_dim, hdim))
        self.q2 = _leafs(nets.init_sale_critic(s, a, zs_dim, hdim))
        self.target_policy = self.policy  # aliased (td7.py:103)
        self.target_q1 = _frozen(self.q1)
        self.target_q2 = _frozen(self.q2)
        self.fixed_encoder = _frozen(self.encoder)
        self.fixed_encoder_target = _frozen(self.encoder)
        self.discount_factor = discount_factor
        self.target_update_rate = target_update_rate
        self.target_policy_noise = target_policy_noise
        self.exploration_noise = exploration_noise
        self.noise_clip = noise_clip
        self.policy_freq = policy_freq
        self.use_lap = use_lap
        self.n_runs = 0
        # td7.py:127-133
        self.optim_policy = torch.optim.Adam(list(self.policy.values()), lr=policy_lr)
        self.optim_q_fns = torch.optim.Adam(list(self.q1.values()) + list(self.q2.values()), lr=critic_lr)
        self.optim_encoder = torch.optim.Adam(list(self.encoder.values()), lr=policy_lr)
        # td7.py:84-87
        self.value_max = -1e8
        self.value_min = 1e8
        self.value_target_max = 0.0
        self.value_target_min = 0.0
        self.last: dict = {}

    # ---- pieces --------------------------------------------------------------
    def encoder_loss(self, state, action, next_state):
        """td7.py:246-257."""
        with torch.no_grad():
            zs_next = nets.encode_state(self.encoder, next_state)
        zs = nets.encode_state(self.encoder, state)
        zsa = nets.encode_state_action(self.encoder, zs, action)
        return (zsa - zs_next).pow(2.0).mean()

    def q_target(self, next_state, reward, done, noise):
        """td7.py:185-218 (no grad)."""
        fet = self.fixed_encoder_target
        zs_n = nets.encode_state(fet, next_state)
        eps = (noise * self.target_policy_noise).clamp(-self.noise_clip, self.noise_clip)
        a_n = (nets.sale_actor(self.target_policy, next_state, zs_n) + eps).clamp(-1.0, 1.0)
        zsa_n = nets.encode_state_action(fet, zs_n, a_n)
        q1n = nets.sale_critic(self.target_q1, next_state, a_n, zsa_n, zs_n)
        q2n = nets.sale_critic(self.target_q2, next_state, a_n, zsa_n, zs_n)
        nv = torch.cat([q1n, q2n], -1).min(1, keepdim=True)[0]
        nv = nv.clamp(self.value_target_min, self.value_target_max)
        return reward + self.discount_factor * nv * done

    def critic_loss(self, state, action, next_state, reward, done, noise):
        """td7.py:175-244."""
        with torch.no_grad():
            y = self.q_target(next_state, reward, done, noise)
            self.value_max = max(self.value_max, float(y.max()))
            self.value_min = min(self.value_min, float(y.min()))
            zs = nets.encode_state(self.fixed_encoder, state)
            zsa = nets.encode_state_action(self.fixed_encoder, zs, action)
        q1 = nets.sale_critic(self.q1, state, action, zsa, zs)
        q2 = nets.sale_critic(self.q2, state, action, zsa, zs)
        self.last.update(q_target=y.detach().clone(), q1=q1.detach().clone(), q2=q2.detach().clone())
        if self.use_lap:
            td = torch.cat([(q1 - y).abs(), (q2 - y).abs()], 1)
            loss = lap_huber_sum(td)
            priority = td.max(1)[0].clamp(1.0).pow(0.4).view(-1)
            return loss, priority.detach()
        loss = torch.mean((y - q1) ** 2.0) * 0.5 + torch.mean((y - q2) ** 2.0) * 0.5
        return loss, None

    def actor_loss(self, state):
        """td7.py:259-276 (fixed_encoder carries no gradient that is ever used, B.4)."""
        with torch.no_grad():
            zs = nets.encode_state(self.fixed_encoder, state)
        a = nets.sale_actor(self.policy, state, zs)
        zsa = nets.encode_state_action(self.fixed_encoder, zs, a)
        q1 = nets.sale_critic(self.q1, state, a, zsa, zs)
        q2 = nets.sale_critic(self.q2, state, a, zsa, zs)
        self.last.update(pi_action=a.detach().clone())
        return -torch.cat([q1, q2], -1).mean()

    def hard_update_target_fns(self):
        """td7.py:279-285, in that order."""
        _copy_into(self.target_q1, self.q1)
        _copy_into(self.target_q2, self.q2)
        _copy_into(self.fixed_encoder_target, self.fixed_encoder)
        _copy_into(self.fixed_encoder, self.encoder)

    # ---- the step ------------------------------------------------------------
    def train_ops(self, batch, replay_buffer=None, noise=None):
        """td7.py:287-332."""
        self.n_runs += 1
        self.last = {}
        info = {}
        s, a, s2, r, d = batch["state"], batch["action"], batch["next_state"], batch["reward"], batch["done"]
        loss_e = self.encoder_loss(s, a, s2)
        self.optim_encoder.zero_grad()
        loss_e.backward()
        self.last["grad_encoder"] = _grads(self.encoder)
        self.optim_encoder.step()
        info["train/encoder"] = float(loss_e.detach())

        if noise is None:
            noise = torch.randn_like(a)
        loss_q, priority = self.critic_loss(s, a, s2, r, d, noise)
        if priority is not None:
            self.last["priority"] = priority.clone()
            replay_buffer.update_priority(priority)
        self.optim_q_fns.zero_grad()
        loss_q.backward()
        self.last["grad_q1"], self.last["grad_q2"] = _grads(self.q1), _grads(self.q2)
        self.optim_q_fns.step()
        info["train/q_fn"] = float(loss_q.detach())

        info["train/policy"] = None
        if self.n_runs % self.policy_freq == 0:
            loss_pi = self.actor_loss(s)
            self.optim_policy.zero_grad()
            loss_pi.backward()
            self.last["grad_policy"] = _grads(self.policy)
            self.optim_policy.step()
            info["train/policy"] = float(loss_pi.detach())
        if self.n_runs % self.target_update_rate == 0:
            self.hard_update_target_fns()
            self.value_target_max = self.value_max
            self.value_target_min = self.value_min
            if self.use_lap:
                replay_buffer.reset_max_priority()
        return info

    @torch.no_grad()
    def sample(self, state, deterministic=False, noise=None):
        """td7.py:141-162."""
        x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
        if x.ndim == 1:
            x = x.unsqueeze(0)
        zs = nets.encode_state(self.fixed_encoder, x)
        act = nets.sale_actor(self.policy, x, zs)
        if not deterministic:
            act = act + (torch.randn_like(act) if noise is None else noise) * self.exploration_noise
        act = np.clip(act.numpy()[0], -1.0, 1.0)
        return act * self.action_scale + self.action_bias

    def nets(self) -> dict:
        return dict(encoder=self.encoder, policy=self.policy, q1=self.q1, q2=self.q2,
                    target_q1=self.target_q1, target_q2=self.target_q2,
                    fixed_encoder=self.fixed_encoder, fixed_encoder_target=self.fixed_encoder_target)


# =============================================================================== TD3
class OracleTD3(_Base):
    """Restatement of ``TD3`` (td3.py:30-245)."""

    def __init__(self, env_id=None, discount_factor=0.99, policy_lr=3e-4, critic_lr=3e-4,
                 exploration_noise=0.1, target_policy_noise=0.2, noise_clip=0.5, policy_freq=2,
                 tau=0.005, use_lap=False, *, state_dim=None, action_dim=None, hidden=256, **_ignored):
        s, a = self._dims(env_id, state_dim, action_dim)
        # td3.py:76-81: policy, q1, q2
        self.policy = _leafs(nets.init_mlp(s, a, hidden))
        self.q1 = _leafs(nets.init_mlp(s + a, 1, hidden))
        self.q2 = _leafs(nets.init_mlp(s + a, 1, hidden))
        self.target_policy = self.policy  # aliased (td3.py:86)
        self.target_q1, self.target_q2 = _frozen(self.q1), _frozen(self.q2)
        self.discount_factor = discount_factor
        self.target_policy_noise = target_policy_noise
        self.exploration_noise = exploration_noise
        self.noise_clip = noise_clip
        self.policy_freq = policy_freq
        self.use_lap = use_lap
        self.tau = tau
        self.n_runs = 0
        self.optim_policy = torch.optim.Adam(list(self.policy.values()), lr=policy_lr)
        self.optim_q_fns = torch.optim.Adam(list(self.q1.values()) + list(self.q2.values()), lr=critic_lr)
        self.last: dict = {}

    def critic_loss(self, s, a, s2, r, d, noise):
        """td3.py:144-182."""
        with torch.no_grad():
            eps = (noise * self.target_policy_noise).clamp(-self.noise_clip, self.noise_clip)
            a2 = (torch.tanh(nets.mlp(self.target_policy, s2)) + eps).clamp(-1.0, 1.0)
            nv = torch.min(nets.mlp_critic(self.target_q1, s2, a2), nets.mlp_critic(self.target_q2, s2, a2))
            y = r + self.discount_factor * nv * d
        q1, q2 = nets.mlp_critic(self.q1, s, a), nets.mlp_critic(self.q2, s, a)
        self.last.update(q_target=y.clone(), q1=q1.detach().clone(), q2=q2.detach().clone())
        if self.use_lap:
            td1, td2 = (q1 - y).abs(), (q2 - y).abs()
            loss = lap_huber_mean(td1) + lap_huber_mean(td2)
            priority = torch.max(td1, td2).clamp(1.0).pow(0.4).view(-1)
            return loss, priority.detach()
        loss = torch.mean((y - q1) ** 2.0) * 0.5 + torch.mean((y - q2) ** 2.0) * 0.5
        return loss, None

    def actor_loss(self, s):
        """td3.py:184-192."""
        a = torch.tanh(nets.mlp(self.policy, s))
        self.last.update(pi_action=a.detach().clone())
        return -torch.min(nets.mlp_critic(self.q1, s, a), nets.mlp_critic(self.q2, s, a)).mean()

    @torch.no_grad()
    def polyak(self):
        """td3.py:195-204 -- q1, q2 and the policy onto itself (aliased target)."""
        for src, dst in ((self.q1, self.target_q1), (self.q2, self.target_q2), (self.policy, self.target_policy)):
            for k in src:
                dst[k].copy_(self.tau * src[k] + dst[k] * (1 - self.tau))

    def _zero(self):
        self.optim_q_fns.zero_grad()
        self.optim_policy.zero_grad()

    def train_ops(self, batch, replay_buffer=None, noise=None):
        """td3.py:206-242."""
        self.last = {}
        info = {}
        s, a, s2, r, d = batch["state"], batch["action"], batch["next_state"], batch["reward"], batch["done"]
        if noise is None:
            noise = torch.randn_like(a)
        loss_q, priority = self.critic_loss(s, a, s2, r, d, noise)
        if priority is not None:
            self.last["priority"] = priority.clone()
            replay_buffer.update_priority(priority)
        loss_q.backward()
        self.last["grad_q1"], self.last["grad_q2"] = _grads(self.q1), _grads(self.q2)
        self.optim_q_fns.step()
        self._zero()
        info["train/q_fn"] = float(loss_q.detach())
        info["train/policy"] = None
        info["norm/policy"] = None
        if self.n_runs % self.policy_freq == 0:
            loss_pi = self.actor_loss(s)
            loss_pi.backward()
            info["train/policy"] = float(loss_pi.detach())
            info["norm/policy"] = grad_norm_sum(self.policy)
            self.last["grad_policy"] = _grads(self.policy)
            self.optim_policy.step()
            self._zero()
            self.polyak()
        self.n_runs += 1
        return info

    @torch.no_grad()
    def sample(self, state, deterministic=False, noise=None):
        """td3.py:114-135."""
        x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
        if x.ndim == 1:
            x = x.unsqueeze(0)
        act = torch.tanh(nets.mlp(self.policy, x))
        if not deterministic:
            act = act + (torch.randn_like(act) if noise is None else noise) * self.exploration_noise
        act = np.clip(act.numpy()[0], -1.0, 1.0)
        return act * self.action_scale + self.action_bias

    def nets(self) -> dict:
        return dict(policy=self.policy, q1=self.q1, q2=self.q2, target_q1=self.target_q1, target_q2=self.target_q2)


# =============================================================================== SAC
_HALF_LOG_2PI = math.log(math.sqrt(2 * math.pi))  # torch/distributions/normal.py log_prob


def gaussian_tanh_sample(mean, log_std, eps):
    """sac.py:154-172 with torch.distributions.Normal spelled out.

    ``x = mean + eps*std``; ``a = tanh(x)``;
    ``log pi = sum(-(x-mean)^2/(2 std^2) - log(std) - log sqrt(2 pi)) - sum log(1 - a^2 + 1e-6)``.
    """
    std = log_std.exp()
    x = mean + eps * std
    a = torch.tanh(x)
    var = std**2
    log_prob = -((x - mean) ** 2) / (2 * var) - std.log() - _HALF_LOG_2PI
    log_pi = log_prob.sum(-1, keepdim=True)
    log_pi = log_pi - torch.log(1 - a.pow(2.0) + EPS).sum(-1, keepdim=True)
    return a, log_pi


class OracleSAC(_Base):
    """Restatement of ``SAC`` (sac.py:24-299), auto-temperature or fixed ``tmp``."""

    def __init__(self, env_id=None, discount_factor=0.99, policy_lr=3e-4, critic_lr=3e-4,
                 min_log_std=-20.0, max_log_std=2.0, tau=0.005, tmp=-1.0, use_lap=False,
                 *, state_dim=None, action_dim=None, hidden=256, **_ignored):
        s, a = self._dims(env_id, state_dim, action_dim)
        # sac.py:80-86: policy (2A outputs), critic01, critic02
        self.policy = _leafs(nets.init_mlp(s, 2 * a, hidden))
        self.q1 = _leafs(nets.init_mlp(s + a, 1, hidden))
        self.q2 = _leafs(nets.init_mlp(s + a, 1, hidden))
        self.target_q1, self.target_q2 = _frozen(self.q1), _frozen(self.q2)
        self.auto_tmp_mode = tmp < 0.0
        self.tmp = torch.zeros(1, requires_grad=True) if self.auto_tmp_mode else tmp
        self.optim_policy = torch.optim.Adam(list(self.policy.values()), lr=policy_lr)
        self.optim_q_fns = torch.optim.Adam(list(self.q1.values()) + list(self.q2.values()), lr=critic_lr)
        self.optim_tmp = torch.optim.Adam([self.tmp], lr=policy_lr) if self.auto_tmp_mode else None
        self.target_entropy = -a
        self.discount_factor = discount_factor
        self.min_log_std, self.max_log_std = min_log_std, max_log_std
        self.tau = tau
        self.n_runs = 0
        self.last: dict = {}

    def _dist(self, p, s):
        mean, log_std = nets.mlp(p, s).chunk(2, -1)
        return mean, torch.clamp(log_std, self.min_log_std, self.max_log_std)

    def _zero(self):
        self.optim_q_fns.zero_grad()
        self.optim_policy.zero_grad()
        if self.auto_tmp_mode:
            self.optim_tmp.zero_grad()

    def critic_loss(self, s, a, s2, r, d, eps):
        """sac.py:174-212 (non-LAP branch; the LAP branch is dead code)."""
        with torch.no_grad():
            a2, logp2 = gaussian_tanh_sample(*self._dist(self.policy, s2), eps)
            nq = torch.min(nets.mlp_critic(self.target_q1, s2, a2), nets.mlp_critic(self.target_q2, s2, a2))
            alpha = self.tmp.exp() if self.auto_tmp_mode else self.tmp
            y = r + self.discount_factor * (nq - alpha * logp2) * d
        q1, q2 = nets.mlp_critic(self.q1, s, a), nets.mlp_critic(self.q2, s, a)
        self.last.update(q_target=y.clone(), q1=q1.detach().clone(), q2=q2.detach().clone())
        return torch.mean((y - q1) ** 2.0) * 0.5 + torch.mean((y - q2) ** 2.0) * 0.5

    def actor_loss(self, s, eps):
        """sac.py:214-241."""
        a, logp = gaussian_tanh_sample(*self._dist(self.policy, s), eps)
        q = torch.min(nets.mlp_critic(self.q1, s, a), nets.mlp_critic(self.q2, s, a))
        alpha = self.tmp.exp().detach() if self.auto_tmp_mode else self.tmp
        policy_obj = torch.mean(-q + logp * alpha)
        tmp_obj = torch.mean(self.tmp.exp() * (-logp.detach() - self.target_entropy)) if self.auto_tmp_mode else 0.0
        self.last.update(pi_action=a.detach().clone(), log_pi=logp.detach().clone())
        return policy_obj, tmp_obj, -(logp.mean().detach())

    @torch.no_grad()
    def polyak(self):
        """sac.py:244-249."""
        for src, dst in ((self.q1, self.target_q1), (self.q2, self.target_q2)):
            for k in src:
                dst[k].copy_(self.tau * src[k] + dst[k] * (1 - self.tau))

    def train_ops(self, batch, replay_buffer=None, noise=None):
        """sac.py:251-295.  ``noise`` = (eps_critic[B,A], eps_actor[B,A])."""
        self.last = {}
        info = {}
        s, a, s2, r, d = batch["state"], batch["action"], batch["next_state"], batch["reward"], batch["done"]
        eps_c = torch.empty_like(a).normal_() if noise is None else noise[0]
        loss_q = self.critic_loss(s, a, s2, r, d, eps_c)
        loss_q.backward()
        self.last["grad_q1"], self.last["grad_q2"] = _grads(self.q1), _grads(self.q2)
        self.optim_q_fns.step()
        self._zero()
        info["train/q_fn"] = float(loss_q.detach())

        eps_a = torch.empty_like(a).normal_() if noise is None else noise[1]
        policy_obj, tmp_loss, entropy = self.actor_loss(s, eps_a)
        obj = policy_obj
        if self.auto_tmp_mode:
            obj += tmp_loss  # in place: policy_obj now also holds the sum (B.8)
        obj.backward()
        self.last["grad_policy"] = _grads(self.policy)
        if self.auto_tmp_mode:
            info["tmp"] = float(self.tmp.exp().detach())
            info["norm/tmp"] = float(self.tmp.grad.data)
        self.optim_policy.step()
        if self.auto_tmp_mode:
            self.optim_tmp.step()
        self._zero()
        info["train/policy"] = float(policy_obj.detach())
        if self.auto_tmp_mode:
            info["train/tmp"] = float(tmp_loss.detach())
        info["entropy"] = float(entropy)
        self.polyak()
        self.n_runs += 1
        return info

    @torch.no_grad()
    def sample(self, state, deterministic=False, noise=None):
        """sac.py:131-152."""
        x = torch.as_tensor(np.asarray(state), dtype=torch.float32)
        if x.ndim == 1:
            x = x.unsqueeze(0)
        mean, log_std = self._dist(self.policy, x)
        if deterministic:
            act = torch.tanh(mean)
        else:
            eps = torch.empty_like(mean).normal_() if noise is None else noise
            act = gaussian_tanh_sample(mean, log_std, eps)[0]
        return act.numpy()[0] * self.action_scale + self.action_bias

    def nets(self) -> dict:
        return dict(policy=self.policy, q1=self.q1, q2=self.q2, target_q1=self.target_q1, target_q2=self.target_q2)
