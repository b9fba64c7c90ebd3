"""A few whole updates of one workload between cudaProfilerStart / Stop, for launch lists and `ncu --set full` captures.

    python profiles/prof_update.py <workload> [updates] [warmup]                      # plain run: CUDA-event time per update
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_<w>_r02.csv python profiles/prof_update.py <workload> 4
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c 1 \
        -o gpurun_out/prof_<kernel>_r02 -f python profiles/prof_update.py <workload> 2
Workloads are bench.py's (td7_ant_pop256, td7_humanoid_pop256, td7_humanoid_1M, td7_halfcheetah_1M, ...)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "td7_ant_pop256"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 6
agent, buf = bench.make_agent_and_buffer(w, 0)
agent.train_ops_n(warm, buf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
agent.train_ops_n(n, buf)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"{w}: {n} updates, {e0.elapsed_time(e1) / n * 1e3:.1f} us per (stacked) update")
