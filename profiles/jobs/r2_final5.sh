cd /root/repo
for w in td7_ant_pop256 td7_humanoid_pop256; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2_ncu_$w.log
done
