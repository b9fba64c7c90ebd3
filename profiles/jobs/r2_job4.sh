python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2_job4_full.log
tail -15 gpurun_out/r2_job4_full.log | cut -c1-300
grep "tcgen05/TF32 parity" gpurun_out/r2_job4_full.log | cut -c1-900
