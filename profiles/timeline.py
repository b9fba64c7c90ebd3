"""Per-table GEMM timeline of one TD7 update (needs a build with -DB200RL_PROBE; GPU only).

    nvcc ... -DB200RL_PROBE ... && B200RL_DEBUG_TABLES=1 python profiles/timeline.py [TD7|TD3|SAC]
Prints, for the LAST update of the burst, every grouped-GEMM launch ordered by start time: table number (the
B200RL_DEBUG_TABLES listing), start relative to the first launch, duration, and the idle gap on the chain.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sac_td3_td7_b200 import _lib  # noqa: E402
from sac_td3_td7_b200.rl.agent import TD7  # noqa: E402
from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory  # noqa: E402

lib = _lib.lib()
algo = sys.argv[1] if len(sys.argv) > 1 else "TD7"
if algo == "TD7":
    env = os.environ.get("TIMELINE_ENV", "HalfCheetah-v4")
    agent = TD7(env, use_lap=True)
    buf = LAPReplayMemory(1_000_000, env)
elif algo == "TD3":
    from sac_td3_td7_b200.rl.agent import TD3
    from sac_td3_td7_b200.rl.replay_memory import SimpleReplayMemory
    env = "Hopper-v4"
    agent = TD3(env)
    buf = SimpleReplayMemory(1_000_000, env)
else:
    from sac_td3_td7_b200.rl.agent import SAC
    from sac_td3_td7_b200.rl.replay_memory import SimpleReplayMemory
    env = "Walker2d-v4"
    agent = SAC(env)
    buf = SimpleReplayMemory(1_000_000, env)
buf.fill_synthetic(1_000_000, seed=0)
agent.train_ops_n(50, buf)
out = (C.c_ulonglong * 512)()
lib.b200rl_debug_timeline.restype = C.c_int
lib.b200rl_debug_timeline(out)
for k in range(4):
    torch.cuda.synchronize()
    agent.train_ops_n(1, buf)
    torch.cuda.synchronize()
    lib.b200rl_debug_timeline(out)
    a = np.array(out[:], dtype=np.uint64).reshape(256, 2)
    rows = [(int(a[t, 0]), int(a[t, 1]), t) for t in range(256) if a[t, 1] > 0]
    rows.sort()
    t0 = rows[0][0]
    print(f"--- update n_runs={agent.n_runs}")
    for s_, e_, t in rows:
        print(f"table {t:3d}  start {(s_ - t0) / 1e3:8.2f} us  end {(e_ - t0) / 1e3:8.2f}  dur {(e_ - s_) / 1e3:6.2f} us")
    print(f"span {(max(r[1] for r in rows) - t0) / 1e3:.1f} us")

# two consecutive updates in ONE burst: per table the earliest entry (first update that used it) and the latest exit
# (last update that used it) -- shows the distance between the last GEMM of one update and the first of the next
torch.cuda.synchronize()
lib.b200rl_debug_timeline(out)
agent.train_ops_n(2, buf)
torch.cuda.synchronize()
lib.b200rl_debug_timeline(out)
a = np.array(out[:], dtype=np.uint64).reshape(256, 2)
rows = [(int(a[t, 0]), int(a[t, 1]), t) for t in range(256) if a[t, 1] > 0]
rows.sort()
t0 = rows[0][0]
print(f"--- burst of two updates ending at n_runs={agent.n_runs}")
for s_, e_, t in rows:
    print(f"table {t:3d}  first start {(s_ - t0) / 1e3:8.2f} us  last end {(e_ - t0) / 1e3:8.2f}")
