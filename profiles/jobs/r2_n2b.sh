cd /root/repo
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r02.json 2> gpurun_out/bench_n${N}_r02.err
echo "rc=$?"; tail -2 gpurun_out/bench_n${N}_r02.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_r02.json').read().strip().splitlines()[-1])
print('N=${N} value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'gathered', d['e2e'].get('gathered_seeds'), d['config']['total_seeds'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
