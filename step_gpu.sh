timeout 600 python -m pytest tests/test_agent_parity.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --workload td3_hopper_1M --steps 3000 --warmup 200 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --workload sac_walker2d_1M --steps 3000 --warmup 200 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
