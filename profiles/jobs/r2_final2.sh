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
    q = float(infos[-1]["train/q_fn"])
    from sac_td3_td7_b200 import _lib
    n = buf.replay_buffer_size
    out = np.empty(n, np.float32)
    ok, mx = np.isfinite(q), 0.0
    for g in (0, 1, 100, 255):  # a NaN anywhere in a seed's critic poisons that seed's priorities
        _lib.check(buf._lib.b200rl_replay_read_priority(buf._h, g, 0, n, out.ctypes.data, _lib.current_stream_ptr()))
        ok = ok and bool(np.isfinite(out).all()) and float(out.max()) < 1e6
        mx = max(mx, float(out.max()))
    print(w, "1500 stacked updates (384 k agent-updates): seed-0 critic loss", q, "| priorities of seeds 0/1/100/255 finite and bounded:", bool(ok), "max", mx)
    assert ok
    del agent, buf
    torch.cuda.empty_cache()
PY
