#!/bin/bash
mkdir -p gpurun_out
bash profiles/run_ncu_full.sh sample_hpop
bash profiles/run_ncu_full.sh sample_apop
tail -3 gpurun_out/ncu_sample_hpop.log
