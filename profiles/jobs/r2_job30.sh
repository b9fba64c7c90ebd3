cd /root/repo
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q 2>&1 | tail -6 | cut -c1-400
