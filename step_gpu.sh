set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_default.log 2>&1; tail -1 gpurun_out/final_default.log | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-300
for w in td3_hopper_1M sac_walker2d_1M td7_humanoid_1M; do timeout 400 python bench.py --workload $w --steps 3000 --warmup 200 > gpurun_out/c4_$w.log 2>&1; tail -1 gpurun_out/c4_$w.log | cut -c1-120; done
