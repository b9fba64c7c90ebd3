timeout 600 python -m pytest tests/test_replay_kernels.py tests/test_population_gpu.py tests/test_rng_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -8
for w in td7_humanoid_pop256 td7_ant_pop256 td7_halfcheetah_1M; do timeout 300 python profiles/prof_sampler.py $w; done
for w in hpop apop; do
  W=td7_humanoid_pop256; [ $w = apop ] && W=td7_ant_pop256
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:lap_sample_kernel -s 8 -c 1 -o gpurun_out/prof_lap_sample_${w}_r02 -f python profiles/prof_sampler.py $W 12 > gpurun_out/ncu_sample_$w.log 2>&1
  tail -2 gpurun_out/ncu_sample_$w.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
