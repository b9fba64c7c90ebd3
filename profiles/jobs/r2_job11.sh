cd /root/repo
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_nccl_gpu.py -m gpu -q -s 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?"; tail -2 gpurun_out/r2_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gathered', d['e2e'].get('gathered_seeds'), d['config']['total_seeds'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-300
