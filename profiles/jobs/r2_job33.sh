cd /root/repo
timeout 900 python -m pytest tests/test_agent_parity.py tests/test_population_gpu.py -m gpu -q 2>&1 | tail -2 | cut -c1-300
for f in 0 1; do for w in td7_humanoid_pop256 td7_ant_pop256; do
  B200RL_TC_FLIP=$f timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1 | sed "s/^/FLIP=$f /"
done; done
B200RL_PDL=1 timeout 300 python profiles/prof_update.py td7_ant_pop256 8 2>&1 | tail -1 | sed "s/^/PDL=1 /"
