cd /root/repo
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default_b.json 2> gpurun_out/r2_bench_default_b.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default_b.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_reference_b.json
cut -c1-400 gpurun_out/r2_bench_reference_b.json
