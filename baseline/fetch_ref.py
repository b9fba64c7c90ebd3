"""Stage the UNMODIFIED reference for bench.py's reference arm.

    python baseline/fetch_ref.py

Copies the pure-Python package ``/root/reference/rl`` to ``baseline/_ref/rl`` (git-ignored, but it travels to the
GPU box with the repo snapshot like the built ``.so``).  ``/root/reference`` only exists in the build container, so this
runs from ``__graft_entry__.build()`` there; on the GPU box ``bench.py --impl reference`` finds the staged copy and times
the reference's own ``LAPReplayMemory.sample`` + ``TD7.train_ops`` loop (``rl/runner/run.py:87-96``) on the host cores,
under the 40-line fake ``gymnasium`` of ``oracle/gym_stub.py`` (MuJoCo is not installed anywhere here).  Nothing is
modified: the copy is byte-identical, and ``pip install`` is not involved because the package has no build step
(``setup.py`` only lists it).  No file of the product package imports it.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/rl"
DST = os.path.join(HERE, "_ref", "rl")


def fetch(verbose: bool = False) -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        cmp = filecmp.dircmp(SRC, DST, ignore=["__pycache__"])
        if not (cmp.left_only or cmp.right_only or cmp.diff_files):
            return DST
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print("staged", SRC, "->", DST)
    return DST


if __name__ == "__main__":
    print(fetch(verbose=True) or "reference not available here", file=sys.stderr if not os.path.isdir(SRC) else sys.stdout)
