#!/usr/bin/env python
"""bench.py -- gradient updates/sec of the off-policy update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

With no --workload the line measures BASELINE.json configs[4] -- the north star's population: TD7 on Ant-v4 shapes,
256 independent seeds stacked per GPU (2 048 on 8 GPUs), 1M-row LAP replay per seed:

* value   : agent-updates/s.  One "step" = one stacked update = every seed of this GPU does ``sample(256)`` +
            ``train_ops`` (priority write-back, amortised actor step / target sync).  K steps through the burst API
            (``train_ops_n``), everything resident in HBM, CUDA events, max over ranks, whole-job aggregate.
* e2e     : the same through the reference-facing Python API with HOST data in the loop: per step one transition per seed
            appended from pinned host memory (H2D, read by the append kernel), ``sample``, ``train_ops``, every seed's
            ``train/q_fn`` read back (D2H), and every 10th step an all-gather of a per-seed metric over NCCL.
* single_agent (N = 1 only): sub-records for ``td7_humanoid_1M`` (the >= 100x target: value, e2e, the CPU arm at 1 thread
            and at all threads, ratio against the better), ``td7_halfcheetah_1M`` (configs[1]), ``sac_walker2d_1M``
            (configs[2]) and ``td7_humanoid_pop256`` (configs[3]).  Each times >= 2 000 updates whatever --steps says
            (a handful of 0.2 ms updates is not a measurement window).
* roofline: the dominant kernel of the stacked update (persistent tcgen05 weight-gradient + Adam launch) against the
            measured HBM peak on SURVEY 8(d)'s algorithmic bytes; ``roofline_replay``: the LAP tree-walk + gather kernel.
* cpu_baseline / --impl reference: the reference's own CPU loop (``rl/runner/run.py:87-96``: ``LAPReplayMemory.sample`` +
            ``TD7.train_ops``) from the staged unmodified copy ``baseline/_ref`` under ``oracle/gym_stub`` when present
            (kind "reference"), else the oracle port (kind "port"); 1 thread and all threads, the better is the value.
            The reference runs seeds one after the other (scripts/td7_exp.sh:1-4), so its agent-updates/s is one agent's.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (algo, env, prioritized, seeds stacked per GPU, buffer rows, one row store shared by the stacked seeds)
    "td7_halfcheetah_1M": ("TD7", "HalfCheetah-v4", True, 1, 1_000_000, False),
    "td7_humanoid_1M": ("TD7", "Humanoid-v4", True, 1, 1_000_000, False),
    "td7_ant_1M": ("TD7", "Ant-v4", True, 1, 1_000_000, False),
    "td3_hopper_1M": ("TD3", "Hopper-v4", False, 1, 1_000_000, False),
    "sac_walker2d_1M": ("SAC", "Walker2d-v4", False, 1, 1_000_000, False),
    # 256 x 1M Ant rows of 256 B = 65.5 GB of private row stores + 1 GB of priorities per GPU
    "td7_ant_pop256": ("TD7", "Ant-v4", True, 256, 1_000_000, False),
    # 256 x 1M Humanoid rows would be 790 GB: one shared 200k-row store (618 MB), private priorities
    "td7_humanoid_pop256": ("TD7", "Humanoid-v4", True, 256, 200_000, True),
}
DEFAULT_WORKLOAD = "td7_ant_pop256"  # BASELINE.json configs[4], the north star's population (256 seeds per GPU)
ENV_DIMS = {"Hopper-v4": (11, 3), "HalfCheetah-v4": (17, 6), "Walker2d-v4": (17, 6), "Ant-v4": (27, 8), "Humanoid-v4": (376, 17)}
BATCH = 256
MIN_SINGLE_UPDATES = 2000
METRIC = "gradient updates/sec (batch 256) per agent and agent-updates/sec at 1/2/4/8 B200"


def algorithmic_gflop(algo: str, S: int, A: int, B: int = BATCH, H: int = 256, Z: int = 256) -> float:
    """Necessary FLOPs of one update (SURVEY.md section 8d formulas; 2 FLOPs per MAC)."""
    if algo == "TD7":
        Es, Esa = S * H + 2 * H * H, (Z + A) * H + 2 * H * H
        PI = S * H + (Z + H) * H + H * H + H * A
        Q = (S + A) * H + (2 * Z + H) * H + H * H + H
        enc = 3 * Es + 2 * Esa + 2 * H * H + (Z * H + 2 * H * H)
        cri = 2 * Es + PI + 2 * Esa + 4 * Q + 2 * Q + 2 * (H + 2 * H * H)
        act = Es + PI + Esa + 2 * Q + 2 * (H + H * H + (H + Z) * H + A * H) + (H * Z + H * H + A * H) + PI + (H * A + 2 * H * H)
        macs = enc + cri + 0.5 * act
    else:
        Aout = 2 * A if algo == "SAC" else A
        PI = S * H + H * H + H * Aout
        Q = (S + A) * H + H * H + H
        cri = PI + 4 * Q + 2 * Q + 2 * (H + H * H)
        act = PI + 2 * Q + 2 * (H + H * H + A * H) + PI + (H * Aout + H * H)
        macs = cri + (act if algo == "SAC" else 0.5 * act)
    return 2.0 * macs * B / 1e9


def algorithmic_mbytes_stacked(S: int, A: int, H: int = 256, Z: int = 256) -> float:
    """SURVEY.md 8(d) "Algorithmic bytes (stacked update)", TD7, per seed-update, fp32: trained parameters with Adam state
    24 B/param (encoder, q1, q2 every update; policy every 2nd), frozen nets read once 4 B/param."""
    enc = (S * H + H) + 2 * (H * H + H) + ((Z + A) * H + H) + 2 * (H * H + H)
    pol = (S * H + H) + ((Z + H) * H + H) + (H * H + H) + (H * A + A)
    q = ((S + A) * H + H) + ((2 * Z + H) * H + H) + (H * H + H) + (H + 1)
    return (24.0 * (enc + 2 * q + 0.5 * pol) + 4.0 * (2 * enc + 2 * q)) / 1e6


def workload_config(name: str, world: int) -> dict:
    algo, env, prio, seeds, rows, share = WORKLOADS[name]
    S, A = ENV_DIMS[env]
    return {"workload": name, "algo": algo, "env": env, "obs_dim": S, "act_dim": A, "batch": BATCH, "buffer_rows": rows,
            "replay": "LAP" if prio else "uniform", "seeds_per_gpu": seeds, "total_seeds": seeds * world,
            "row_store": "shared by the stacked seeds (private priorities)" if share else "private per seed",
            "parallelism": f"seeds sharded over {world} GPU(s), no data-path collective; NCCL all-gather of per-seed metrics only",
            "l2": "replay rows and, for stacked seeds, weights + Adam state exceed L2 (inputs larger than L2)"}


# ------------------------------------------------------------------------------------------------ CPU arm
def _ref_dir():
    d = os.path.join(ROOT, "baseline", "_ref")
    return d if os.path.isdir(os.path.join(d, "rl")) else None


def _make_cpu_loop(workload: str):
    """(step_fn, kind): one CPU update of the workload's single agent -- the unmodified reference when staged, else the port."""
    import numpy as np
    import torch

    algo, env, prio, _, rows, _ = WORKLOADS[workload]
    ref = _ref_dir()
    torch.manual_seed(0)
    if ref is not None:
        from oracle import gym_stub

        gym_stub.import_reference(ref)
        from rl.agent import SAC, TD3, TD7  # noqa: PLC0415  (the reference's own classes)
        from rl.replay_memory import LAPReplayMemory, SimpleReplayMemory  # noqa: PLC0415

        agent = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo](env, **({} if algo == "SAC" else {"use_lap": prio}))
        buf = (LAPReplayMemory if prio else SimpleReplayMemory)(rows, env)
        g = np.random.default_rng(0)  # BASELINE.md section 2 synthetic contents, written straight into the reference's arrays
        for lo in range(0, rows, 65536):
            hi = min(rows, lo + 65536)
            buf.state[lo:hi] = g.standard_normal((hi - lo, buf.state.shape[1]), dtype=np.float32)
            buf.action[lo:hi] = g.random((hi - lo, buf.action.shape[1]), dtype=np.float32) * 2.0 - 1.0
            buf.reward[lo:hi] = g.standard_normal((hi - lo, 1), dtype=np.float32)
            buf.next_state[lo:hi] = g.standard_normal((hi - lo, buf.state.shape[1]), dtype=np.float32)
            buf.float_done[lo:hi] = (g.random((hi - lo, 1), dtype=np.float32) < 0.99)
        if prio:
            buf.priority[:rows] = 1.0
        buf.size, buf.ptr = rows, 0
        kind = "reference"
    else:
        from oracle import agents as oagents
        from oracle import replay as oreplay

        agent = {"TD7": oagents.OracleTD7, "TD3": oagents.OracleTD3, "SAC": oagents.OracleSAC}[algo](env, use_lap=prio)
        buf = (oreplay.OracleLAP if prio else oreplay.OracleUniform)(rows, env)
        buf.fill_synthetic(rows, seed=0)
        kind = "port"

    def step():
        agent.train_ops(buf.sample(BATCH), replay_buffer=buf)  # rl/runner/run.py:92-95

    return step, kind


def cpu_updates_per_s(workload: str, budget_s: float, updates: int | None = None, warmup: int = 5, thread_settings=None):
    """Times the CPU loop at each thread setting; returns {"value", "cores", "kind", "sample", "by_threads"}: value = the
    better setting ("1 thread and all threads, take the better")."""
    import torch

    ncpu = max(1, os.cpu_count() or 1)
    settings = thread_settings or ([1, ncpu] if ncpu > 1 else [1])
    step, kind = _make_cpu_loop(workload)
    by, best = {}, None
    for th in settings:
        torch.set_num_threads(th)  # (also under torchrun, which exports OMP_NUM_THREADS=1 to its workers)
        for _ in range(warmup):
            step()
        n, t0 = 0, time.perf_counter()
        while True:
            step()
            n += 1
            el = time.perf_counter() - t0
            if (updates is not None and n >= updates) or (updates is None and el >= budget_s / len(settings)) or el > 120:
                break
        by[str(th)] = {"updates_per_s": n / el, "updates": n, "seconds": el}
        if best is None or n / el > best[0]:
            best = (n / el, th, n, el)
    ups, th, n, el = best
    return {"value": ups, "unit": "updates/s", "cores": th, "kind": kind, "by_threads": by,
            "sample": f"{n} updates of {'the unmodified reference (baseline/_ref, fake gymnasium)' if kind == 'reference' else 'the oracle port of the reference loop'}"
                      f" on {workload}'s single agent, torch {torch.__version__} CPU, best of threads {settings} = {th} (os.cpu_count={ncpu}), {el:.1f}s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # A reference-arm step is a bounded sample of the workload: `per_step` consecutive updates of one agent, sized so that the
    # K steps hold a few hundred updates per thread setting and the whole run stays within a few minutes.
    K, W = args.steps, args.warmup
    per_step = max(1, -(-300 // K))
    cpu = cpu_updates_per_s(args.workload, budget_s=40.0, updates=K * per_step if args.steps_given else None, warmup=max(W, 5))
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "updates/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "updates_per_step": per_step if args.steps_given else None,
        "ms_per_step": 1e3 * per_step / cpu["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference trains the seeds of a population one after the other (scripts/td7_exp.sh:1-4): its agent-updates/s is one agent's updates/s",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    FIELDS = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self, timed=None):
        """Median SM clock and the throttle reasons seen inside `timed` = (t0, t1) wall-clock bounds (the timed regions);
        when they hold no sample (shorter than the 100 ms sampling period) the whole loaded window is reported and says so."""
        import datetime

        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = -1.0
            try:
                rows.append((ts, float(parts[1]), float(parts[2]), [nm for nm, v in zip(names, parts[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        window = "whole loaded window (warm-up + timed regions)"
        if timed is not None:
            inside = [r for r in rows if timed[0] <= r[0] <= timed[1]]
            if inside:
                rows, window = inside, "timed regions"
        sm = sorted(r[1] for r in rows)
        mx = rows[-1][2] if rows else None
        reasons = {nm for r in rows for nm in r[3]}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def make_agent_and_buffer(workload: str, rank: int):
    import torch

    from sac_td3_td7_b200.rl.agent import SAC, TD3, TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    algo, env, prio, seeds, rows, share = WORKLOADS[workload]
    torch.manual_seed(1000 + rank)
    cls = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo]
    agent = cls(env, use_lap=prio, n_seeds=seeds) if algo != "SAC" else cls(env, n_seeds=seeds)
    if seeds > 1:
        agent.seed_population(1000 * (rank + 1))
    buf = (LAPReplayMemory if prio else SimpleReplayMemory)(rows, env, n_seeds=seeds, share_rows=share)
    buf.fill_synthetic(rows, seed=rank)
    return agent, buf


def timed_burst(agent, buf, K, W, dist=None):
    """K updates through train_ops_n, CUDA events on the launching stream, barrier + synchronize on both sides."""
    import torch

    stream = torch.cuda.current_stream()
    agent.train_ops_n(W, buf)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.time()
    e0.record(stream)
    infos = agent.train_ops_n(K, buf)
    e1.record(stream)
    torch.cuda.synchronize()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms, infos, (w0, w1)


def e2e_loop(agent, buf, steps, W, world, total_seeds, dist=None):
    """The reference-facing per-step API with host data in the loop (rl/runner/run.py:160-171 per environment step)."""
    import numpy as np
    import torch

    from sac_td3_td7_b200.population import gather_per_seed

    G, S, A = agent.n_seeds, agent.state_dim, agent.action_dim
    g = np.random.default_rng(5)
    obs = g.standard_normal((64, G, S), dtype=np.float32)
    act = ((g.random((64, G, A), dtype=np.float32) * 2 - 1) * np.asarray(agent.action_scale, np.float32)).astype(np.float32)  # environment units
    rew = g.standard_normal((64, G), dtype=np.float32)
    nxt = g.standard_normal((64, G, S), dtype=np.float32)
    nd = np.ones((64, G), np.float32)
    col = agent.INFO_KEYS.index("train/q_fn")
    gathered = None

    def step(i):
        nonlocal gathered
        j = i % 64
        if G == 1:
            buf.append([obs[j, 0], act[j, 0], rew[j, 0], nxt[j, 0], nd[j, 0]])  # one transition: pinned staging slot, read by the append kernel
            info = agent.train_ops(buf.sample(BATCH), replay_buffer=buf)
            q = float(info["train/q_fn"])  # D2H: the loss mailbox of this update
            if dist is not None and i % 10 == 9:
                gathered = gather_per_seed(torch.tensor([q], device="cuda"), total_seeds)
            return q
        buf.append_population(obs[j], act[j], rew[j], nxt[j], nd[j])  # one transition per seed, one launch
        agent.train_ops(buf.sample(BATCH), replay_buffer=buf)
        q = agent.read_infos(agent._n_calls, 1, -1)[:, 0, col]  # every seed's loss: one strided D2H + sync
        if dist is not None and i % 10 == 9:  # metrics all-gather over NCCL / NVLink (eval returns in a real run)
            gathered = gather_per_seed(torch.from_numpy(q).cuda(), total_seeds)
        return q

    for i in range(W):
        step(i)
    if dist is not None:  # (the first collective of a shape sets up its NCCL channels: part of the warm-up, like the graph capture)
        gather_per_seed(torch.zeros(G, device="cuda"), total_seeds)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([el], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    cols = len(agent.INFO_KEYS)
    rs = buf.row_stride
    return {"value": world * G * steps / el, "unit": "updates/s", "h2d_bytes_per_step": (1 if buf.share_rows else G) * rs * 4, "d2h_bytes_per_step": G * cols * 4 if G > 1 else 8,
            "steps": steps, "seconds": el, "gathered_seeds": None if gathered is None else int(gathered.shape[0]),
            "note": ("per step: one transition per seed appended from the library's pinned staging ring (read over the bus by the append kernel), sample(256) "
                     "straight into the engine's input buffers, train_ops, every seed's train/q_fn read back; every 10th step an NCCL all-gather of the per-seed "
                     "metric across ranks" if G > 1 else
                     "per step: append 1 transition from pinned host memory, sample(256), train_ops, float(info['train/q_fn']) from the pinned loss mailbox")}


def act_append_latency(agent, buf):
    """Host-to-host latency of agent.sample(obs) and of replay.append(transition), microseconds."""
    import numpy as np
    import torch

    G, S, A = agent.n_seeds, agent.state_dim, agent.action_dim
    g = np.random.default_rng(1)
    obs = g.standard_normal((G, S)).astype(np.float32)
    o = obs if G > 1 else obs[0]
    torch.cuda.synchronize()
    for _ in range(20):
        agent.sample(o, True)
    t0 = time.perf_counter()
    n = 300
    for _ in range(n):
        agent.sample(o, True)
    act_us = (time.perf_counter() - t0) / n * 1e6
    # the same call at the C ABI (no Python argument handling), and the kernel alone (device tensors, CUDA events)
    from sac_td3_td7_b200 import _lib

    out = np.empty((G, agent._act_out_dim), np.float32)
    po, pout, st = obs.ctypes.data, out.ctypes.data, _lib.current_stream_ptr()
    t0 = time.perf_counter()
    for _ in range(n):
        agent._lib.b200rl_agent_act_host(agent._h, po, pout, st)
    act_c_us = (time.perf_counter() - t0) / n * 1e6
    d_obs, d_out = torch.as_tensor(obs).cuda(), torch.empty((G, agent._act_out_dim), device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        agent.policy_out_device(d_obs, d_out)
    torch.cuda.synchronize()
    e0.record(torch.cuda.current_stream())
    for _ in range(n):
        agent._lib.b200rl_agent_act(agent._h, d_obs.data_ptr(), d_out.data_ptr(), st)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    act_kernel_us = e0.elapsed_time(e1) / n * 1e3
    a = (g.random((G, A), dtype=np.float32) * 2 - 1) * np.asarray(agent.action_scale, np.float32)
    r, nd = np.zeros(G, np.float32), np.ones(G, np.float32)
    ap = (lambda: buf.append([obs[0], a[0], 0.0, obs[0], 1.0])) if G == 1 else (lambda: buf.append_population(obs, a, r, obs, nd))
    for _ in range(20):
        ap()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        ap()
    host_us = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    total_us = (time.perf_counter() - t0) / n * 1e6
    return {"act_us": act_us, "act_c_abi_us": act_c_us, "act_kernel_us": act_kernel_us, "append_host_us": host_us, "append_us": total_us,
            "note": "act_us: agent.sample(obs, deterministic) host to host (pinned observation slot -> one cluster kernel per seed -> pinned action mailbox, no stream "
                    "sync); act_c_abi_us: b200rl_agent_act_host alone; act_kernel_us: back-to-back launches of the acting kernel on device tensors; append_host_us: host time of replay.append per call (asynchronous); append_us: the same including the device work (synchronised at the end)"}


def single_agent_record(workload: str, K: int, with_cpu: bool, lib):
    import torch

    algo, env, prio, seeds, rows, share = WORKLOADS[workload]
    S, A = ENV_DIMS[env]
    agent, buf = make_agent_and_buffer(workload, 0)
    K = max(K, MIN_SINGLE_UPDATES) if seeds == 1 else max(K, 20)
    ms, infos, _ = timed_burst(agent, buf, K, 300 if seeds == 1 else 3)  # (single agents: long enough for the SM clock to ramp)
    value = seeds * K / (ms / 1e3)
    e2e = e2e_loop(agent, buf, min(K, 2000) if seeds == 1 else min(K, 50), 5, 1, seeds)
    lat = act_append_latency(agent, buf)
    lp0, lp1 = lib.b200rl_agent_launches_per_update(agent._h, 0), lib.b200rl_agent_launches_per_update(agent._h, 1)
    rec = {"workload": workload, "config": workload_config(workload, 1), "value": value, "unit": "updates/s", "us_per_update": 1e3 * ms / K if seeds == 1 else None,
           "ms_per_stacked_update": ms / K if seeds > 1 else None, "timed_updates": K, "e2e": e2e, **lat,
           "launches_per_update": {"critic_only": lp0, "with_actor": lp1}, "algorithmic_gflop_per_update": algorithmic_gflop(algo, S, A),
           "last_info": {k: (None if v is None else float(v)) for k, v in infos[-1].items()}}
    if seeds > 1:
        rec["roofline_whole_update"] = whole_update_roofline(value, S, A)
    del agent, buf
    torch.cuda.empty_cache()
    if with_cpu:
        cpu = cpu_updates_per_s(workload if seeds == 1 else workload.replace("pop256", "1M"), budget_s=24.0)
        rec["cpu_baseline"] = cpu
        rec["speedup_vs_cpu"] = {"device_resident": value / cpu["value"], "e2e": e2e["value"] / cpu["value"],
                                 "against": f"the better of the CPU arm's thread settings ({cpu['cores']} threads, kind {cpu['kind']})"}
    return rec


def whole_update_roofline(agent_updates_per_s_per_gpu: float, S: int, A: int):
    hbm_peak, _, _, src = measured_peaks()
    mb = algorithmic_mbytes_stacked(S, A)
    ach = mb * 1e6 * agent_updates_per_s_per_gpu / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "algorithmic_mbytes_per_seed_update": mb, "peak_source": src,
            "note": "SURVEY 8(d) algorithmic bytes of one TD7 seed-update (weights + Adam state of the trained nets, frozen nets once) x agent-updates/s per GPU"}


def stacked_roofline(lib, torch, G, mode):
    """Graph-timed launch of the tcgen05 grouped GEMM on `G` stacked seeds: two 256x256x256 layers (mode 0: forward +
    bias + ELU; mode 2: weight gradient + fused Adam).  Returns (flops, algorithmic bytes, ms per launch)."""
    from sac_td3_td7_b200 import _lib

    dev = torch.device("cuda")
    Bn, H, K = 256, 256, 256
    keep, probs = [], []
    for _ in range(2):
        p = _lib.GemmProblem()
        if mode == 0:
            x, w = torch.randn(G, Bn, K, device=dev), torch.randn(G, H, K, device=dev) / 16
            b, y = torch.randn(G, H, device=dev), torch.empty(G, Bn, H, device=dev)
            p.I, p.J, p.C, p.c_ld, p.c_gs, p.act, p.bias, p.bias_gs = Bn, H, y.data_ptr(), H, Bn * H, 2, b.data_ptr(), H
            sg = p.seg[0]
            sg.A, sg.a_ld, sg.a_gs, sg.B, sg.b_ld, sg.b_gs, sg.len = x.data_ptr(), K, Bn * K, w.data_ptr(), K, H * K, K
            keep += [x, w, b, y]
        else:
            dz, x = torch.randn(G, Bn, H, device=dev) / Bn, torch.randn(G, Bn, K, device=dev)
            w, m, v = torch.randn(G, H, K, device=dev) / 16, torch.zeros(G, H, K, device=dev), torch.zeros(G, H, K, device=dev)
            sc = _lib.AdamScalars(w1=0.1, beta2=0.999, w2=0.001, bc2_sqrt=0.0316, neg_step=-3e-3, eps=1e-8, step=1)
            scd = torch.frombuffer(bytearray(bytes(sc)), dtype=torch.uint8).to(dev)
            p.I, p.J, p.c_ld = H, K, K
            p.P, p.M, p.V, p.p_ld, p.p_gs, p.adam = w.data_ptr(), m.data_ptr(), v.data_ptr(), K, H * K, scd.data_ptr()
            sg = p.seg[0]
            sg.A, sg.a_ld, sg.a_gs, sg.B, sg.b_ld, sg.b_gs, sg.len = dz.data_ptr(), H, Bn * H, x.data_ptr(), K, Bn * K, Bn
            keep += [dz, x, w, m, v, scd]
        p.nseg = 1
        probs.append(p)
    arr = (_lib.GemmProblem * 2)(*probs)
    ms = C.c_float()
    _lib.check(lib.b200rl_op_gemm_timed(mode, arr, 2, G, 3, 20, C.byref(ms), torch.cuda.current_stream().cuda_stream))
    flops = 2 * 2.0 * G * Bn * H * K
    byts = 2 * G * 4.0 * (Bn * K + H * K + Bn * H) if mode == 0 else 2 * G * 4.0 * (Bn * H + Bn * K + 6 * H * K)
    return flops, byts, ms.value


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from sac_td3_td7_b200 import _lib, build

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the update path has no CPU fallback)")
    torch.cuda.set_device(local)
    D = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        D = dist
    if rank == 0:
        build.build()
    if D:
        D.barrier()
    lib = _lib.lib()

    algo, env, prio, seeds, rows, share = WORKLOADS[args.workload]
    S, A = ENV_DIMS[env]
    K, W = args.steps, max(args.warmup, 3)
    if seeds == 1 and not args.steps_given:
        K = MIN_SINGLE_UPDATES
    agent, buf = make_agent_and_buffer(args.workload, rank)

    clocks = ClockSampler(local) if rank == 0 else None
    ms_max, infos, (w0, w1) = timed_burst(agent, buf, K, W, D)
    value = world * seeds * K / (ms_max / 1e3)
    last = {k: (None if v is None else float(v)) for k, v in infos[-1].items()}
    e2e_steps = min(K, 2000) if seeds == 1 else min(K, 100)
    e2e = e2e_loop(agent, buf, e2e_steps, W, world, seeds * world, D)
    w2 = time.time()
    clk = clocks.stop(timed=(w0, w2)) if clocks else None
    lat = act_append_latency(agent, buf) if rank == 0 else None
    lp0, lp1 = lib.b200rl_agent_launches_per_update(agent._h, 0), lib.b200rl_agent_launches_per_update(agent._h, 1)
    if rank != 0:
        if D:
            D.barrier()
            D.destroy_process_group()
        return

    # ---- rooflines (rank 0, standalone graph-timed launches) ----
    hbm_peak, bf16_burst, bf16_sus, peak_src = measured_peaks()
    roofline, roof_extra = None, {}
    if seeds >= 8:
        # dominant kernel of the stacked update (35 % of its kernel time, profiles/launches_td7_ant_pop256_r02.summary.txt): persistent tcgen05 weight gradient + fused Adam
        fl, by, ms3 = stacked_roofline(lib, torch, seeds, 2)
        ach = by / (ms3 * 1e-3) / 1e9
        roofline = {"kernel": f"gemm_tc_ws_kernel<wgrad> (persistent tcgen05 TF32, TMA-fed, fused Adam), 2 layers x {seeds} seeds", "bound": "hbm", "achieved": ach,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": 1009.6e6, "launch_ms": ms3, "bytes_per_launch": by,
                    "peak_source": peak_src + " (burst: kernel timed alone)",
                    "note": "algorithmic bytes per launch = 2 layers x seeds x 4 B x (dZ 256x256 + X 256x256 + p, m, v in and out 6 x 256x256), SURVEY 8(d) 24 B/param; "
                            "traffic = dram__bytes_read + write of the same launch, profiles/gemm_tc_ws_mode2_ncu_full_r02.summary.txt (658.7 MB read + 350.9 MB written)",
                    "tensor": {"achieved": fl / (ms3 * 1e-3) / 1e12, "peak": bf16_sus, "unit": "TFLOP/s", "frac": fl / (ms3 * 1e-3) / 1e12 / bf16_sus}}
        fl0, by0, ms0 = stacked_roofline(lib, torch, seeds, 0)
        roof_extra["roofline_forward"] = {"kernel": "gemm_tc_ws_kernel<fwd>, 2 layers x %d seeds" % seeds, "bound": "hbm", "achieved": by0 / (ms0 * 1e-3) / 1e9, "peak": hbm_peak,
                                          "unit": "GB/s", "frac": by0 / (ms0 * 1e-3) / 1e9 / hbm_peak, "launch_ms": ms0, "bytes_per_launch": by0,
                                          "tensor": {"achieved": fl0 / (ms0 * 1e-3) / 1e12, "peak": bf16_sus, "unit": "TFLOP/s", "frac": fl0 / (ms0 * 1e-3) / 1e12 / bf16_sus}}
        roof_extra["roofline_whole_update"] = whole_update_roofline(value / world, S, A)
    roof2 = None
    if prio:
        ms2 = (C.c_float * 2)()
        _lib.check(lib.b200rl_replay_time_sample(buf._h, BATCH, 10, 100, ms2, torch.cuda.current_stream().cuda_stream))
        gather_bytes = seeds * BATCH * 2.0 * 4.0 * (2 * S + A + 2)  # rows read + batch tensors written
        roof2 = {"kernel": "lap_sample_kernel (3-level fp64 sum-tree walk + row gather, one warp per draw)", "bound": "hbm", "achieved": gather_bytes / (ms2[1] * 1e-3) / 1e9,
                 "peak": hbm_peak, "unit": "GB/s", "frac": gather_bytes / (ms2[1] * 1e-3) / 1e9 / hbm_peak, "launch_ms": ms2[1], "bytes_per_launch": gather_bytes,
                 "peak_source": peak_src,
                 "note": "algorithmic bytes = seeds x B x 2 x 4(2S+A+2) (SURVEY 8(d) gather term; the tree levels a draw reads are L2-resident and not counted)",
                 "full_rebuild_kernel": {"kernel": "lap_tile_sums_kernel (bulk loads / restored checkpoints only)", "launch_ms": ms2[0], "bytes_per_launch": 4.0 * rows * seeds,
                                         "achieved_gbs": 4.0 * rows * seeds / (ms2[0] * 1e-3) / 1e9}}

    # ---- CPU baseline on this host (rank 0, N = 1), and the single-agent sub-records ----
    cpu, singles = None, None
    del agent, buf
    torch.cuda.empty_cache()
    # (every GPU measurement first, the CPU arms last: a 16-thread torch run leaves its worker pool spinning beside the
    # thread that launches the next workload -- the launch-bound SAC record read 155 us per update behind it, 123 us alone)
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_sub_records:
        singles = {}
        for name in ("td7_humanoid_1M", "td7_halfcheetah_1M", "sac_walker2d_1M", "td7_humanoid_pop256"):
            singles[name] = single_agent_record(name, K, with_cpu=False, lib=lib)
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_updates_per_s(args.workload if seeds == 1 else args.workload.replace("pop256", "1M"), budget_s=24.0)
        if singles and "td7_humanoid_1M" in singles:
            rec = singles["td7_humanoid_1M"]
            c1 = cpu_updates_per_s("td7_humanoid_1M", budget_s=24.0)
            rec["cpu_baseline"] = c1
            rec["speedup_vs_cpu"] = {"device_resident": rec["value"] / c1["value"], "e2e": rec["e2e"]["value"] / c1["value"],
                                     "against": f"the better of the CPU arm's thread settings ({c1['cores']} threads, kind {c1['kind']})"}

    per_update = (lp0 + lp1) / 2.0 if algo != "SAC" else float(lp1)
    gflop = algorithmic_gflop(algo, S, A)
    line = {
        "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (stacked seeds: TF32 tensor-core inputs, fp32 accumulate and state)" if seeds >= 8 else "f32",
        "data": "synthetic", "config": workload_config(args.workload, world),
        "e2e": e2e, "gpu_launches": int(K * (per_update + 1)), "launches_per_update": {"critic_only": lp0, "with_actor": lp1},
        "clocks": clk, "roofline": roofline, "roofline_replay": roof2, **roof_extra, "cpu_baseline": cpu, **(lat or {}),
        "algorithmic_gflop_per_update": gflop, "update_tflops_per_gpu": gflop * value / 1e3 / world,
        "single_agent": singles, "last_info": last,
    }
    if roofline is None:  # single-agent workloads: the launch-bound update has no bandwidth / tensor roofline; report the FFMA fraction
        line["roofline"] = {"kernel": "whole update (grouped fp32 FFMA launches replayed as one CUDA graph)", "bound": "tensor", "achieved": gflop * value / 1e3 / world, "peak": bf16_burst,
                            "unit": "TFLOP/s", "frac": gflop * value / 1e3 / world / bf16_burst, "traffic": None, "peak_source": peak_src,
                            "note": "fp32 FFMA by design (1e-5 fp32 parity); the non-tensor fp32 peak of a B200 is ~74.5 TFLOP/s",
                            "frac_of_fp32_ffma_peak": gflop * value / 1e3 / world / 74.5}
    if cpu:
        line["speedup_vs_cpu"] = {"device_resident": value / cpu["value"], "e2e": e2e["value"] / cpu["value"],
                                  "against": f"one agent on the host CPU, better of the thread settings ({cpu['cores']} threads, kind {cpu['kind']})"}
    print(json.dumps(line), flush=True)
    if D:
        D.barrier()
        D.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 50
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
