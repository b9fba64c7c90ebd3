cd /root/repo
nvidia-smi -L | head -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
timeout 600 python -m pytest tests/test_nccl_gpu.py -m gpu -q -s 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err
echo "rc=$?"; tail -2 gpurun_out/bench_n2_r02.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r02.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gathered', d['e2e'].get('gathered_seeds'), d['config']['total_seeds'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference_n2_r02.json
cut -c1-250 gpurun_out/bench_reference_n2_r02.json
