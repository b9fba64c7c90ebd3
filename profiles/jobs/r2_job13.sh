cd /root/repo
timeout 300 python - <<'PY' 2>&1 | grep -v "^\[b200rl\]" | tail -8
import sys; sys.path.insert(0,'.')
import bench, json
for w in ("td7_humanoid_1M","td7_halfcheetah_1M","td3_hopper_1M"):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    agent.train_ops_n(50, buf)
    print(w, json.dumps({k:(round(v,2) if isinstance(v,float) else v) for k,v in bench.act_append_latency(agent, buf).items() if k!='note'}))
PY
B200RL_DEBUG_TABLES=1 timeout 300 python profiles/prof_update.py td7_ant_pop256 2 > gpurun_out/r2_tables_ant_pop256.txt 2>&1
grep -c "gemm table" gpurun_out/r2_tables_ant_pop256.txt
