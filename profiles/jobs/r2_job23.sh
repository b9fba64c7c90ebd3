cd /root/repo
for d in 0 2 4 6 7; do
  B200RL_NW_DBG=$d timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:narrow_wgrad -c 4 --csv --log-file gpurun_out/nw_dbg$d.csv python profiles/prof_update.py td7_ant_pop256 1 2 > /dev/null 2>&1
  echo "dbg=$d: $(python profiles/summarize_launches.py gpurun_out/nw_dbg$d.csv 10 | grep "^('" | tr '\n' ' ')"
done
