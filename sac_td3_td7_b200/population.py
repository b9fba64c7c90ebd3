"""Seed populations across GPUs: one process per GPU, seeds sharded, no data-path collective.

The per-agent update does not shard, so independent seeds are the parallel
unit: every rank owns ``count`` stacked seeds (weights, Adam state, priorities stay on its GPU) and the
only communication is an ``all_gather`` of per-seed evaluation returns / metrics (~1 KB per rank),
NCCL over NVLink on GPUs, gloo in the CPU tests.  The reference runs seeds as separate sequential
processes (scripts/td7_exp.sh:1-4).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_seeds(total_seeds: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first_seed, count) of `rank`: contiguous blocks, the remainder spread over the first ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_seeds, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_per_seed(local: torch.Tensor, total_seeds: int, group=None) -> torch.Tensor:
    """All ranks' per-seed values in global seed order: local [count, ...] -> [total_seeds, ...].

    Ranks may own different counts (remainder), so every rank pads to the maximum block size and the
    padding is dropped after the gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.clone()
    world = dist.get_world_size(group)
    max_count = -(-total_seeds // world)
    pad = torch.zeros((max_count,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * max_count,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        _, c = shard_seeds(total_seeds, world, r)
        parts.append(out[r * max_count: r * max_count + c])
    return torch.cat(parts, 0)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing convention of bench.py: a multi-GPU time is the max over ranks."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class Population:
    """`count` stacked TD7 / TD3 / SAC seeds of this rank plus their (shared-row) replay memory."""

    def __init__(self, agent_cls, replay_cls, env_id: str, total_seeds: int, base_seed: int = 0, buffer_rows: int = 1_000_000,
                 rank: int | None = None, world_size: int | None = None, **agent_kwargs):
        self.rank = dist.get_rank() if rank is None and dist.is_initialized() else (rank or 0)
        self.world = dist.get_world_size() if world_size is None and dist.is_initialized() else (world_size or 1)
        self.total_seeds = total_seeds
        self.first, self.count = shard_seeds(total_seeds, self.world, self.rank)
        self.agent = agent_cls(env_id, n_seeds=self.count, **agent_kwargs)
        self.agent.seed_population(base_seed + self.first)
        self.replay = replay_cls(buffer_rows, env_id, n_seeds=self.count, share_rows=True)

    def train(self, n_updates: int):
        return self.agent.train_ops_n(n_updates, self.replay)

    def evaluate(self, envs, n_episodes: int = 10, deterministic: bool = True):
        """``test_agent`` (rl/runner/run.py:21-45) for every stacked seed of this rank at once: seed g acts in
        ``envs[g]`` (gymnasium API: ``reset() -> (obs, info)``, ``step(a) -> (obs, r, terminated, truncated, info)``);
        each environment step is ONE batched acting call over all seeds (agents whose episode has ended act on their last
        observation and the action is dropped).  Returns the mean undiscounted return per seed, numpy [count]."""
        import numpy as np

        G = self.count
        assert len(envs) == G, "one environment per stacked seed"
        totals = np.zeros(G, np.float64)
        for _ in range(n_episodes):
            obs = np.stack([np.asarray(envs[g].reset()[0], np.float32) for g in range(G)])
            live = np.ones(G, bool)
            while live.any():
                act = np.asarray(self.agent.sample(obs if G > 1 else obs[0], deterministic)).reshape(G, -1)
                for g in range(G):
                    if not live[g]:
                        continue
                    o, r, terminated, truncated, _ = envs[g].step(act[g])
                    totals[g] += float(r)
                    obs[g] = np.asarray(o, np.float32)
                    live[g] = not (terminated or truncated)
        return totals / n_episodes

    def gather_returns(self, local_returns: torch.Tensor) -> torch.Tensor:
        """Per-seed evaluation returns of the whole population in global seed order (NCCL all-gather, ~1 KB per rank)."""
        if not torch.is_tensor(local_returns):
            local_returns = torch.as_tensor(local_returns, dtype=torch.float32)
        if dist.is_initialized() and dist.get_backend() == "nccl":
            local_returns = local_returns.to(self.agent.device)
        return gather_per_seed(local_returns, self.total_seeds)
