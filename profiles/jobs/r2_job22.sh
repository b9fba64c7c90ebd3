cd /root/repo
timeout 900 python -m pytest tests/test_agent_parity.py tests/test_population_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
for w in sac_walker2d_1M sac_walker2d_1M td3_hopper_1M; do timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['launches_per_update'], d['clocks'])
"; done
for w in td7_ant_pop256 td7_humanoid_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
w=td7_ant_pop256
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02d.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_${w}_r02d.csv | head -14
