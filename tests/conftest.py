import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    """C-ABI driver: 'cuda' = the product library on a GPU (parity proper); 'emu' = host emulation
    of the same kernel sources (logic pre-check in the GPU-less container)."""
    import devlib

    d = devlib.Dev(request.param)
    yield d
    d.free_all()


@pytest.fixture(params=["mma", "ffma"])
def gemm_path(request, dev):
    """Every GEMM-family test runs on both product loops of the small-tile kernel: 3xTF32 on the tensor cores (the
    default) and fp32 FFMA.  The host emulation has only the FFMA loop."""
    if dev.backend == "emu":
        if request.param == "mma":
            pytest.skip("the host emulation runs the FFMA loop only")
        yield
        return
    dev.lib.b200rl_set_mma(1 if request.param == "mma" else 0)
    yield
    dev.lib.b200rl_set_mma(-1)
