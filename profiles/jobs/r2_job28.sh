cd /root/repo
timeout 900 python -m pytest tests/test_agent_parity.py tests/test_population_gpu.py tests/test_tc_gemm_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
for w in td7_humanoid_pop256 td7_ant_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
w=td7_ant_pop256
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02e.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_${w}_r02e.csv 2>/dev/null | head -8
