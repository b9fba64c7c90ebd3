python profiles/jobs/diag_parity.py mma td3_ant_lap td7_halfcheetah_lap > gpurun_out/r2_diag_mma.log 2>&1
python profiles/jobs/diag_parity.py ffma td3_ant_lap > gpurun_out/r2_diag_ffma.log 2>&1
grep -c . gpurun_out/r2_diag_mma.log
