cd /root/repo
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_agent_parity.py tests/test_population_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -3 | cut -c1-300
for w in td7_humanoid_pop256 td7_ant_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
