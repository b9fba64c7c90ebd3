# ncu --set full capture of ONE kernel per call (pass: ffma | tc0 | tc2 | sample)
set -e
case "$1" in
  ffma)
    export B200_GEMM_WARMUP=4 B200_GEMM_ITERS=4
    python profiles/prof_gemm.py > gpurun_out/prof_ffma_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/prof_gemm_ffma_r01 -f python profiles/prof_gemm.py > gpurun_out/prof_ffma_ncu.log 2>&1
    tail -1 gpurun_out/prof_ffma_plain.log ;;
  tc0|tc2)
    export ITERS=4 MODE=${1#tc}
    python profiles/prof_gemm_tc.py > gpurun_out/prof_tc${MODE}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_gemm_tc_mode${MODE}_r01 -f python profiles/prof_gemm_tc.py > gpurun_out/prof_tc${MODE}_ncu.log 2>&1
    tail -1 gpurun_out/prof_tc${MODE}_plain.log ;;
  sample)
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lap_sample_kernel -s 10 -c 1 -o gpurun_out/prof_lap_sample_r01 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
    tail -1 gpurun_out/plain.log | cut -c1-200 ;;
esac
# persistent warp-specialised tcgen05 kernel: bash profiles/run_ncu_full.sh ws0 | ws2
case "$1" in
  ws0|ws2)
    export ITERS=4 MODE=${1#ws}
    python profiles/prof_gemm_tc.py > gpurun_out/prof_ws${MODE}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ws -s 3 -c 1 -o gpurun_out/prof_gemm_tc_ws_mode${MODE}_r01 -f python profiles/prof_gemm_tc.py > gpurun_out/prof_ws${MODE}_ncu.log 2>&1
    tail -1 gpurun_out/prof_ws${MODE}_plain.log ;;
esac
# stacked-seed LAP sampler (256 seeds x 256 draws in one launch): bash profiles/run_ncu_full.sh sample_hpop | sample_apop
case "$1" in
  sample_hpop|sample_apop)
    W=td7_humanoid_pop256; [ "$1" = sample_apop ] && W=td7_ant_pop256
    python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lap_sample_kernel -s 6 -c 1 -o gpurun_out/prof_lap_${1}_r01 -f python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
    tail -1 gpurun_out/plain_$1.log | cut -c1-200 ;;
esac
