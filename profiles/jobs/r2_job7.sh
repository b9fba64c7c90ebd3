timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('seconds'))
print('roofline', d['roofline']['frac'], d.get('roofline_whole_update',{}).get('frac'), 'replay', d['roofline_replay']['frac'])
print('act/append', d.get('act_us'), d.get('append_host_us'), d.get('append_us'))
print('cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d.get('speedup_vs_cpu'))
for k,v in (d.get('single_agent') or {}).items():
    print(k, 'value', round(v['value'],1), 'us', v.get('us_per_update'), 'e2e', round(v['e2e']['value'],1), 'act', round(v['act_us'],1), 'append', round(v['append_host_us'],1), round(v['append_us'],1), v.get('speedup_vs_cpu'), (v.get('cpu_baseline') or {}).get('by_threads'))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-600
