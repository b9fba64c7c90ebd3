"""Phase timeline inside one FFMA GEMM CTA (needs a -DB200RL_PROBE build; GPU only): SM-clock cycles block (0,0)
spends in setup / prologue issue / first wait / main loop / reduction / epilogue, and the gap between launches."""
import ctypes as C
import os
import sys

import numpy as np

sys.argv = [sys.argv[0]]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep_gemm.py")).read().split('print("mode nprob')[0]
exec(src)
for (mode, nprob, K) in [(0, 1, 64), (0, 1, 256), (0, 5, 256), (1, 1, 256), (2, 1, 256), (2, 5, 256)]:
    us = launch(mode, nprob, K, 1)
    out = (C.c_longlong * 640)()
    lib.b200rl_debug_probe.restype = C.c_int
    lib.b200rl_debug_probe(out)
    a = np.array(out[:]).reshape(64, 10)
    a = a[np.argsort(a[:, 0])]
    gaps = a[1:, 0] - a[:-1, 9]
    dur = a[:, 9] - a[:, 0]
    cyc = a[:, 2:9] - a[:, 1:8]
    print(f"mode {mode} nprob {nprob} K {K}: {us:.2f} us/launch | block0 ns median {np.median(dur):.0f}, gap to next launch ns median {np.median(gaps):.0f}")
    print("   cycles (setup, prologue issue, first wait, main loop, reduce, epilogue math, store):", np.median(cyc, axis=0).astype(int))
