cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-sub-records 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],1), d['launches_per_update'], d['clocks'])
"
