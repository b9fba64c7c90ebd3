"""CPU-side checks of the C ABI: the nvcc-built library loads, exports every symbol the header
declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200rl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from sac_td3_td7_b200 import _lib, build

    build.build()
    cdll = _lib.lib()
    names = header_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/b200rl.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert cdll.b200rl_is_cuda_build() == 1
    assert cdll.b200rl_version() >= 100


def test_struct_layouts_match_the_header():
    from sac_td3_td7_b200 import _lib

    # b200rl_gemm_seg: 2 pointers, 2 int64, 4 int32 ; b200rl_adam_scalars: 6 float + 2 int
    assert C.sizeof(_lib.GemmSeg) == 48 and C.sizeof(_lib.AdamScalars) == 32
    assert C.sizeof(_lib.GemmProblem) % 8 == 0 and _lib.GemmProblem.seg.offset == 0
    assert _lib.GemmProblem.nseg.offset == 3 * 48


def test_no_device_means_loud_failure_not_cpu_fallback():
    import torch

    from sac_td3_td7_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cdll = _lib.lib()
    h = C.c_void_p()
    st = cdll.b200rl_replay_create(1000, 11, 3, 1, 1, 0, C.byref(h))
    assert st == -3 and b"no CPU fallback" in cdll.b200rl_last_error()
    from sac_td3_td7_b200.rl.agent import TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory

    with pytest.raises(_lib.B200RLError):
        LAPReplayMemory(1000, "Hopper-v4")
    with pytest.raises(_lib.B200RLError):
        TD7("Hopper-v4")


def test_package_never_imports_the_oracle_or_the_emulator():
    pkg = os.path.join(ROOT, "sac_td3_td7_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                if f.endswith(".py"):
                    assert "cuemu" not in src and "build_emu" not in src, f
