cd /root/repo
timeout 1200 python -m pytest tests/test_agent_parity.py tests/test_population_gpu.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
for d in 0 2 4; do
  B200RL_NW_DBG=$d timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:narrow_wgrad -c 3 --csv --log-file gpurun_out/nw_dbg$d.csv python profiles/prof_update.py td7_ant_pop256 1 2 > /dev/null 2>&1
  echo "dbg=$d: $(python profiles/summarize_launches.py gpurun_out/nw_dbg$d.csv 10 | grep "^('" | tr '\n' ' ')"
done
for w in td7_ant_pop256 td7_humanoid_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
