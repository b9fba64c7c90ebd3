"""The stacked LAP sampler alone (256 seeds x 256 draws per launch), for timing and `ncu -k regex:lap_sample_kernel`.

    python profiles/prof_sampler.py td7_humanoid_pop256 | td7_ant_pop256 | td7_halfcheetah_1M [launches]
Prints the CUDA-event time per launch and the algorithmic gather rate (seeds x B x 2 x 4(2S+A+2) bytes)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from sac_td3_td7_b200 import _lib  # noqa: E402
from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "td7_humanoid_pop256"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
algo, env, prio, seeds, rows, share = bench.WORKLOADS[w]
S, A = bench.ENV_DIMS[env]
buf = LAPReplayMemory(rows, env, n_seeds=seeds, share_rows=share)
buf.fill_synthetic(rows, seed=0)
lib = _lib.lib()
ms = (C.c_float * 2)()
_lib.check(lib.b200rl_replay_time_sample(buf._h, 256, 5, iters, ms, torch.cuda.current_stream().cuda_stream))
by = seeds * 256 * 2.0 * 4.0 * (2 * S + A + 2)
print(f"{w}: lap_sample_kernel {ms[1] * 1e3:.1f} us per launch, gather {by / 1e6:.1f} MB -> {by / (ms[1] * 1e-3) / 1e9:.0f} GB/s algorithmic "
      f"({by / (ms[1] * 1e-3) / 1e9 / 6543.1:.2f} of the measured HBM copy peak); row stride {buf.row_stride * 4} B")
