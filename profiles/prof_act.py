"""The acting kernel alone: a few agent.sample() calls between cudaProfilerStart / Stop (ncu -k regex:act_fused)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "td7_halfcheetah_1M"
agent, buf = bench.make_agent_and_buffer(w, 0)
agent.train_ops_n(10, buf)
obs = np.random.default_rng(0).standard_normal(agent.state_dim).astype(np.float32)
for _ in range(5):
    agent.sample(obs, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(5):
    agent.sample(obs, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(w, "ok")
