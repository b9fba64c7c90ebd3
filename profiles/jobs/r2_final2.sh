cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2_final2_tests.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_r02.json 2> gpurun_out/bench_default_r02.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_default_r02.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference_r02.json
cut -c1-200 gpurun_out/bench_reference_r02.json
# long-run sanity of the stacked path: 1500 stacked updates (384 k agent-updates), every seed's losses finite, sum tree exact
timeout 900 python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, bench
for w in ("td7_ant_pop256", "td7_humanoid_pop256"):
    agent, buf = bench.make_agent_and_buffer(w, 0)
    infos = agent.train_ops_n(1500, buf)
    torch.cuda.synchronize()
    last = agent.info_dicts(agent.n_runs - 8, 8) if hasattr(agent, "info_dicts") else infos[-8:]
    vals = np.array([[float(v) for v in d.values() if v is not None] for d in (last if isinstance(last, list) else [last])], dtype=np.float64)
    pr = buf.priority
    ok = np.isfinite(vals).all() and bool(torch.isfinite(pr).all())
    print(w, "1500 stacked updates: losses finite", bool(np.isfinite(vals).all()), "priorities finite", bool(torch.isfinite(pr).all()), "max_priority", float(torch.as_tensor(buf.max_priority).max()), "q_fn tail", vals[-1][:3])
    assert ok
    del agent, buf
    torch.cuda.empty_cache()
PY
