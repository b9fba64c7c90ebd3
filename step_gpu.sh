timeout 600 python -m pytest tests/test_agent_parity.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 3000 --warmup 200 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
