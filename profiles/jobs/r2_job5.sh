timeout 600 python -m pytest tests/test_agent_parity.py -m gpu -q -x -k "ffma and td7 and not tcgen05" 2>&1 | tail -30 > gpurun_out/r2_job5_tests.log
tail -5 gpurun_out/r2_job5_tests.log
for m in 1 0; do for w in td7_halfcheetah_1M td7_humanoid_1M; do B200RL_MEGA=$m timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('MEGA=$m', d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['launches_per_update'], d['last_info'])
"; done; done > gpurun_out/r2_job5_bench.log 2>&1
cat gpurun_out/r2_job5_bench.log
B200RL_DEBUG_TABLES=1 timeout 120 python -c "
import time,torch
from sac_td3_td7_b200.rl.agent import TD7
t0=time.time(); a=TD7('HalfCheetah-v4', use_lap=True); print('create s', time.time()-t0)
" 2>&1 | grep -E "mega|create"
