"""Standalone launcher of the grouped GEMM kernel for ncu (a few launches, nothing else).

    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 2 -o gpurun_out/prof python profiles/prof_gemm.py [mode]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from sac_td3_td7_b200 import _lib  # noqa: E402

lib = _lib.lib()
torch.cuda.set_device(0)
S, A = 17, 6
flops, ms = bench.gemm_roofline(lib, torch, S, A)
print(f"gemm launch: {ms*1e3:.2f} us, {flops/ms/1e9:.2f} TFLOP/s")
