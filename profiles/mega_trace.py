"""Per-tile trace of the persistent dataflow kernels of one TD7 update (needs the -DB200RL_PROBE build; GPU only).

    python -m sac_td3_td7_b200.build --probe
    B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so python profiles/mega_trace.py [HalfCheetah-v4]

For every program (A / B of the critic-only and the actor variant, C) of the LAST update: span, tiles, and per tile kind the
median time between taking the ticket and having the dependencies (wait), computing (run) and publishing (signal)."""
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

env = sys.argv[1] if len(sys.argv) > 1 else "HalfCheetah-v4"
lib = _lib.lib()
agent = TD7(env, use_lap=True)
buf = LAPReplayMemory(1_000_000, env)
buf.fill_synthetic(1_000_000, seed=0)
agent.train_ops_n(50, buf)
out = np.zeros((8, 4096, 5), np.uint64)
lib.b200rl_debug_mega_trace.restype = C.c_int
lib.b200rl_debug_mega_trace(out.ctypes.data_as(C.c_void_p))
KINDS = ["fwd", "dgrad", "wgrad", "norm", "norm_bwd", "prep", "gauss"]
for rep in range(2):
    torch.cuda.synchronize()
    agent.train_ops_n(1, buf)
    torch.cuda.synchronize()
    lib.b200rl_debug_mega_trace(out.ctypes.data_as(C.c_void_p))
    print(f"=== update n_runs={agent.n_runs}")
    t00 = None
    for p in range(8):
        a = out[p]
        live = a[:, 3] > 0
        if not live.any():
            continue
        a = a[live].astype(np.int64)
        t0, t1 = a[:, 0].min(), a[:, 3].max()
        t00 = t0 if t00 is None else min(t00, t0)
        kind = (a[:, 4] >> 32).astype(int)
        sm = (a[:, 4] & 0xfff).astype(int)
        print(f"program {p}: {len(a)} tiles on {len(set(sm))} SMs, start {(t0 - t00) / 1e3:7.1f} us, span {(t1 - t0) / 1e3:6.1f} us; busy fraction of CTA time "
              f"{(a[:, 2] - a[:, 1]).sum() / max(1, (a[:, 3] - a[:, 0]).sum()):.2f}")
        for k in sorted(set(kind)):
            m = kind == k
            w, r, s = a[m, 1] - a[m, 0], a[m, 2] - a[m, 1], a[m, 3] - a[m, 2]
            print(f"    {KINDS[k]:9s} n={m.sum():5d}  wait med {np.median(w) / 1e3:6.2f} p90 {np.percentile(w, 90) / 1e3:6.2f} | run med {np.median(r) / 1e3:6.2f} p90 {np.percentile(r, 90) / 1e3:6.2f} max {r.max() / 1e3:6.2f}"
                  f" | signal med {np.median(s) / 1e3:5.2f} us")
        # tiles in flight over time (10 buckets)
        edges = np.linspace(t0, t1, 11)
        run = [(((a[:, 1] < edges[i + 1]) & (a[:, 2] > edges[i])).sum()) for i in range(10)]
        print("    tiles computing per tenth of the span:", run)
