cd /root/repo
for pf in 1 0; do for w in td7_halfcheetah_1M td3_hopper_1M; do
  B200RL_ACT_PREFETCH=$pf timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:act_fused --csv --log-file gpurun_out/act_$pf.csv python profiles/prof_act.py $w > /dev/null 2>&1
  echo "prefetch=$pf $w: $(python profiles/summarize_launches.py gpurun_out/act_$pf.csv 10 | grep "^('" | tr '\n' ' ')"
done; done
for w in td7_humanoid_1M td7_halfcheetah_1M; do timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['launches_per_update'])
"; done
timeout 600 python -m pytest tests/test_agent_parity.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
