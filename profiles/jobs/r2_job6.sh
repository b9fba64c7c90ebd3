B200RL_LIB=$PWD/sac_td3_td7_b200/libb200rl_probe.so timeout 300 python profiles/mega_trace.py > gpurun_out/r2_mega_trace.txt 2>&1
tail -60 gpurun_out/r2_mega_trace.txt
