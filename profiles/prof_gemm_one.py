"""ncu target: ONE grouped-GEMM problem (256x256xK fp32, 64 tiles) launched a few times.

    NPROB=1 K=768 MODE=0 B200RL_SPLITK=1 ncu --set full --warp-sampling-interval 0 --import-source on -k regex:gemm_kernel -s 4 -c 1 ... python profiles/prof_gemm_one.py
"""
import os
import sys

sys.argv = [sys.argv[0]]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep_gemm.py")).read().split('print("mode nprob')[0]
exec(src)
mode, nprob, K = int(os.environ.get("MODE", 0)), int(os.environ.get("NPROB", 1)), int(os.environ.get("K", 768))
sk = int(os.environ.get("B200RL_SPLITK", 1))
print(f"mode {mode} nprob {nprob} K {K} splitk {sk}: {launch(mode, nprob, K, sk):.2f} us per launch")
