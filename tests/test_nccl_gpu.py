"""Population metrics over NCCL (SURVEY 8e "Collective"): ``gather_per_seed`` all-gathers per-seed evaluation returns /
metrics across one process per GPU.  Needs >= 2 visible GPUs (``gpurun --gpus 2``); the same function is covered with gloo
on CPU in tests/test_population_gloo.py, and bench.py runs it inside its timed end-to-end loop at N > 1."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gather_per_seed_over_nccl_two_ranks():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tests", "_nccl_worker.py"), ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and res.stdout.count("OK") == 2, res.stdout[-3000:] + res.stderr[-6000:]
