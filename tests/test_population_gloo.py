"""N > 1 host logic on CPU: two gloo ranks shard the seeds and all-gather per-seed metrics."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sac_td3_td7_b200.population import gather_per_seed, max_over_ranks, shard_seeds


def test_shard_seeds_partitions_exactly():
    for total in (1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            got = [shard_seeds(total, world, r) for r in range(world)]
            assert sum(c for _, c in got) == total
            nxt = 0
            for first, count in got:
                assert first == nxt
                nxt += count
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    assert shard_seeds(2048, 8, 3) == (768, 256)  # BASELINE.json config 5: 256 seeds per GPU


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_seeds(total, world, rank)
    local = torch.arange(first, first + count, dtype=torch.float32).unsqueeze(1) * torch.tensor([[1.0, 10.0]])
    allv = gather_per_seed(local, total)
    t = max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, allv.tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 8])
def test_all_gather_of_per_seed_metrics_two_ranks(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), 10.0 * i] for i in range(total)]
    for rank, allv, t in res:
        assert allv == want, rank          # global seed order on every rank, remainder handled
        assert t == 2.0                     # max over ranks
