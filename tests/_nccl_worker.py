"""Worker of tests/test_nccl_gpu.py (one process per GPU under torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, sys.argv[1])
from sac_td3_td7_b200.population import Population, shard_seeds  # noqa: E402
from sac_td3_td7_b200.rl.agent import TD7  # noqa: E402
from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
TOTAL = 5  # uneven shards: 3 + 2
pop = Population(TD7, LAPReplayMemory, "Hopper-v4", total_seeds=TOTAL, base_seed=7, buffer_rows=4096, use_lap=True)
assert (pop.first, pop.count) == shard_seeds(TOTAL, world, rank)
pop.replay.fill_synthetic(4096, seed=1)
pop.train(6)
torch.cuda.synchronize()
q = pop.agent.read_infos(6, 1, -1)[:, 0, 1]  # train/q_fn of every local seed after 6 updates
local_vals = torch.tensor(np.asarray(q, np.float32).reshape(pop.count), device="cuda")
allv = pop.gather_returns(local_vals)  # NCCL all_gather_into_tensor, global seed order
assert allv.shape == (TOTAL,) and allv.is_cuda and torch.isfinite(allv).all()
assert torch.equal(allv[pop.first: pop.first + pop.count], local_vals)
ref = allv.clone()  # every rank holds the same vector
dist.broadcast(ref, 0)
assert torch.equal(ref, allv)
if rank == 0:
    print("gathered", [round(float(x), 5) for x in allv.cpu()])
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
