"""Build ``libb200rl.so`` in-tree with nvcc for sm_100a.

    python -m sac_td3_td7_b200.build [--force]

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting ``.so``
is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200rl.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
]


def sources() -> list[str]:
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "b200rl.h"))
    return out


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(s) <= t for s in sources())


def build_probe() -> str:
    """Profiling build with the in-kernel timeline (-DB200RL_PROBE): libb200rl_probe.so, loaded with B200RL_LIB=..."""
    out = os.path.join(HERE, "libb200rl_probe.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    res = subprocess.run([nvcc, *NVCC_FLAGS, "-DB200RL_PROBE", os.path.join(CSRC, "b200rl.cu"), "-o", out], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building the probe library")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libb200rl.so")
    cmd = [nvcc, *NVCC_FLAGS, os.path.join(CSRC, "b200rl.cu"), "-o", OUT + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libb200rl.so")
    if verbose:
        print(res.stderr)
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    if "--probe" in sys.argv:
        print(build_probe())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
