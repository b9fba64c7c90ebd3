cd /root/repo
for m in 0 1; do for w in td7_humanoid_1M td7_halfcheetah_1M sac_walker2d_1M; do B200RL_MMA=$m timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('MMA=$m', d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1))
"; done; done
