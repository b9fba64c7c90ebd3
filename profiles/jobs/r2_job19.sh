cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
for pf in 1 0; do
B200RL_ACT_PREFETCH=$pf timeout 300 python - <<'PY' 2>&1 | grep -v "^\[b200rl\]" | tail -4
import sys, os; sys.path.insert(0,'.')
import bench, json
for w in ("td7_humanoid_1M","td7_halfcheetah_1M","td3_hopper_1M"):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    agent.train_ops_n(50, buf)
    print("prefetch", os.environ["B200RL_ACT_PREFETCH"], w, json.dumps({k:(round(v,2) if isinstance(v,float) else v) for k,v in bench.act_append_latency(agent, buf).items() if k!='note'}))
PY
done
for w in td7_humanoid_1M td7_halfcheetah_1M; do timeout 300 python bench.py --workload $w --steps 2000 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['config']['workload'], 'value', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1))
"; done
