python -m pytest tests/test_population_gpu.py tests/test_rng_gpu.py -m gpu -q -x -s 2>&1 | tail -60 > gpurun_out/r2_job3_newtests.log
python -m pytest tests/test_agent_parity.py -m gpu -q -s -k "tcgen05 or (mma and td7_halfcheetah_lap)" 2>&1 | tail -150 > gpurun_out/r2_job3_tc.log
B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so B200RL_MMA=1 B200RL_DEBUG_TABLES=1 python profiles/timeline.py TD7 > gpurun_out/r2_timeline_mma1.txt 2>&1
B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so B200RL_MMA=0 python profiles/timeline.py TD7 > gpurun_out/r2_timeline_mma0.txt 2>&1
B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so B200RL_MMA=1 python profiles/probe_phases.py > gpurun_out/r2_phases_mma1.txt 2>&1
tail -5 gpurun_out/r2_job3_newtests.log; tail -5 gpurun_out/r2_job3_tc.log
