cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
for w in td7_ant_pop256 td7_humanoid_pop256 td7_humanoid_1M td7_halfcheetah_1M; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
w=td7_ant_pop256
B200RL_DEBUG_TABLES=1 timeout 300 python profiles/prof_update.py $w 1 2 2>&1 | grep -E "TMA" | sort | uniq -c | sort -rn | head
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02c.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_${w}_r02c.csv | head -22
