cd /root/repo
for d in 0 1 2 3 4 7; do
B200RL_ACT_DBG=$d timeout 300 python - <<'PY' 2>&1 | grep -v "^\[b200rl\]" | tail -2
import sys, os, subprocess, threading, time; sys.path.insert(0,'.')
import bench, json
clk=[]
stop=False
def samp():
    while not stop:
        try: clk.append(int(subprocess.run(["nvidia-smi","--query-gpu=clocks.sm","--format=csv,noheader,nounits"],capture_output=True,text=True).stdout.split()[0]))
        except Exception: pass
        time.sleep(0.05)
for w in ("td7_halfcheetah_1M",):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    agent.train_ops_n(50, buf)
    th=threading.Thread(target=samp); th.start()
    r=bench.act_append_latency(agent, buf)
    stop=True; th.join()
    print("dbg", os.environ["B200RL_ACT_DBG"], w, {k:round(v,2) for k,v in r.items() if k.startswith('act')}, "sm clocks during the probe", sorted(set(clk)))
PY
done
