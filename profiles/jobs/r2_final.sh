cd /root/repo
nvidia-smi -L | head -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2_final_tests.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_r02.json 2> gpurun_out/bench_default_r02.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_default_r02.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference_r02.json
cut -c1-200 gpurun_out/bench_reference_r02.json
for w in td7_ant_pop256 td7_humanoid_pop256 td7_humanoid_1M td7_halfcheetah_1M; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_r02.csv python profiles/prof_update.py $w 4 > gpurun_out/r2_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2_ncu_$w.log
done
# ncu --set full of the dominant stacked kernels (one launch each), after the plain run above exited 0
w=td7_ant_pop256
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_ws_kernel -s 20 -c 1 -o gpurun_out/prof_stacked_k1_r02 -f python profiles/prof_update.py $w 2 > gpurun_out/r2_ncufull1.log 2>&1
for m in 2 0 1; do
  MODE=$m ITERS=4 timeout 300 python profiles/prof_gemm_tc.py > gpurun_out/prof_ws${m}_plain_r02.log 2>&1 && MODE=$m ITERS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ws -s 3 -c 1 -o gpurun_out/prof_gemm_tc_ws_mode${m}_r02 -f python profiles/prof_gemm_tc.py > gpurun_out/prof_ws${m}_ncu_r02.log 2>&1
  tail -1 gpurun_out/prof_ws${m}_plain_r02.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lap_sample -s 8 -c 1 -o gpurun_out/prof_lap_sample_apop_r02 -f python profiles/prof_sampler.py td7_ant_pop256 12 > gpurun_out/ncu_sample_apop.log 2>&1
tail -1 gpurun_out/ncu_sample_apop.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:narrow_wgrad -s 1 -c 1 -o gpurun_out/prof_narrow_wgrad_r02 -f python profiles/prof_update.py $w 1 2 > gpurun_out/r2_ncufull2.log 2>&1
# timeline of single TD7 updates (probe build)
B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so B200RL_DEBUG_TABLES=1 timeout 300 python profiles/timeline.py TD7 > gpurun_out/timeline_td7_r02.txt 2>&1
grep -c "table" gpurun_out/timeline_td7_r02.txt; grep span gpurun_out/timeline_td7_r02.txt | head -4
ls -la gpurun_out/*r02*.ncu-rep
