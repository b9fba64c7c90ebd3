cd /root/repo
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x -k "narrow or rank1 or non_tileable or stacked_groups" -p no:cacheprovider > gpurun_out/memcheck_narrow_r02.log 2>&1
echo "memcheck narrow rc=$?"; tail -4 gpurun_out/memcheck_narrow_r02.log | cut -c1-200
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/prof_update.py td7_ant_pop256 2 2 > gpurun_out/memcheck_pop_r02.log 2>&1
echo "memcheck pop rc=$?"; tail -3 gpurun_out/memcheck_pop_r02.log | cut -c1-200
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x -k "narrow or rank1" -p no:cacheprovider > gpurun_out/racecheck_narrow_r02.log 2>&1
echo "racecheck narrow rc=$?"; tail -4 gpurun_out/racecheck_narrow_r02.log | cut -c1-200
