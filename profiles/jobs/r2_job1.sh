python -m pytest tests/test_op_kernels.py tests/test_agent_parity.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_job1_tests.log
for m in 1 0; do for w in td7_halfcheetah_1M td7_humanoid_1M; do B200RL_MMA=$m python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('MMA=$m', d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'gemm_ms', d['roofline']['launch_ms'])
"; done; done > gpurun_out/r2_job1_bench.log 2>&1
cat gpurun_out/r2_job1_tests.log gpurun_out/r2_job1_bench.log
