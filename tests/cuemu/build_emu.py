"""Build the TEST-ONLY host emulation of libb200rl (g++ -DB200_EMU, see cuemu.h).

The emulated library exists so kernel indexing / synchronisation / orchestration logic can be
exercised on the GPU-less build container.  It is never loaded by the product package.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sac_td3_td7_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libb200rl_emu.so")


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuemu.h"), os.path.join(ROOT, "include", "b200rl.h")]
    return all(os.path.getmtime(s) <= t for s in srcs)


def build(force: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-DB200_EMU", "-x", "c++", f"-I{HERE}", "-mfma", "-mavx2", "-ffp-contract=off",
           "-fPIC", "-shared", os.path.join(CSRC, "b200rl.cu"), "-o", OUT + ".tmp", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulation library")
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
