cd /root/repo
timeout 900 python -m pytest tests/test_narrow_kernels.py tests/test_tc_gemm_gpu.py -m gpu -q 2>&1 | tail -3 | cut -c1-300
for w in td7_humanoid_pop256 td7_ant_pop256; do timeout 300 python profiles/prof_update.py $w 8 2>&1 | tail -1; done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:narrow_wgrad -c 4 --csv --log-file gpurun_out/nw_dbg0.csv python profiles/prof_update.py td7_ant_pop256 1 2 > /dev/null 2>&1
echo "nw: $(python profiles/summarize_launches.py gpurun_out/nw_dbg0.csv 10 | grep "^('" | tr '\n' ' ')"
