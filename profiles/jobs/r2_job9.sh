timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250
for w in td7_ant_pop256 td7_humanoid_pop256; do timeout 300 python profiles/prof_sampler.py $w; done
B200RL_SAMPLE_SUB=0 timeout 300 python profiles/prof_sampler.py td7_ant_pop256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lap_sample -s 8 -c 1 -o gpurun_out/prof_lap_sample_apop_r02 -f python profiles/prof_sampler.py td7_ant_pop256 12 > gpurun_out/ncu_sample_apop.log 2>&1
tail -1 gpurun_out/ncu_sample_apop.log
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench, json
from sac_td3_td7_b200 import _lib
for w in ("td7_humanoid_1M","td3_hopper_1M"):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    agent.train_ops_n(50, buf)
    print(w, json.dumps({k:(round(v,2) if isinstance(v,float) else v) for k,v in bench.act_append_latency(agent, buf).items() if k!='note'}))
PY
