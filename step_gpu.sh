set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in td7_halfcheetah_1M td7_humanoid_1M td3_hopper_1M sac_walker2d_1M; do
  timeout 400 python bench.py --workload $w --steps 3000 --warmup 200 > gpurun_out/c2_$w.log 2>&1; tail -1 gpurun_out/c2_$w.log | cut -c1-200
done
timeout 400 python bench.py --workload td7_ant_pop256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c2_pop256.log 2>&1; tail -1 gpurun_out/c2_pop256.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash profiles/run_ncu_pop.sh
