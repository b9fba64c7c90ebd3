cd /root/repo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench, json
for w in ("td7_humanoid_1M","td3_hopper_1M"):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    agent.train_ops_n(50, buf)
    print(w, json.dumps({k:(round(v,2) if isinstance(v,float) else v) for k,v in bench.act_append_latency(agent, buf).items() if k!='note'}))
PY
