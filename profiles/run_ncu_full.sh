# ncu --set full captures: current FFMA grouped GEMM (single agent) and the tcgen05 grouped GEMM (256 stacked seeds)
export B200_GEMM_WARMUP=4 B200_GEMM_ITERS=4
python profiles/prof_gemm.py > gpurun_out/prof_ffma_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/prof_gemm_ffma_r01 -f python profiles/prof_gemm.py > gpurun_out/prof_ffma_ncu.log 2>&1
tail -1 gpurun_out/prof_ffma_plain.log
export ITERS=4
for MODE in 0 2; do
  export MODE
  python profiles/prof_gemm_tc.py > gpurun_out/prof_tc${MODE}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_gemm_tc_mode${MODE}_r01 -f python profiles/prof_gemm_tc.py > gpurun_out/prof_tc${MODE}_ncu.log 2>&1
  tail -1 gpurun_out/prof_tc${MODE}_plain.log
done
ITERS=50 MODE=0 python profiles/prof_gemm_tc.py; ITERS=50 MODE=2 python profiles/prof_gemm_tc.py
