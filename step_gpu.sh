timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_replay_kernels.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python profiles/e2e_prof.py 2>&1 | tail -3
