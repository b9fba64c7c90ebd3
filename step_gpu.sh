timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 3000 --warmup 200 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --workload td7_ant_pop256 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
