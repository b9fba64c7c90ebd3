cd /root/repo
nvidia-smi -L | head -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_job12_tests.log
# mega A/B with the constant-bank descriptor table + double-buffered tile records
for m in 1 0; do for w in td7_halfcheetah_1M td7_humanoid_1M; do B200RL_MEGA=$m timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('MEGA=$m', d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['launches_per_update'])
"; done; done > gpurun_out/r2_job12_mega.log 2>&1
cat gpurun_out/r2_job12_mega.log
# r02 launch lists
for w in td7_ant_pop256 td7_humanoid_1M; do
  timeout 300 python profiles/prof_update.py $w 4 > gpurun_out/r2_plain_$w.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2_plain_$w.log; tail -1 gpurun_out/r2_ncu_$w.log
done
# timeline (probe build) of single TD7 Humanoid updates
B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so B200RL_DEBUG_TABLES=1 TIMELINE_ENV=Humanoid-v4 timeout 300 python profiles/timeline.py TD7 > gpurun_out/r2_timeline_humanoid.txt 2>&1
tail -3 gpurun_out/r2_timeline_humanoid.txt
