#!/usr/bin/env python
"""bench.py -- gradient updates/sec of the off-policy update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one update = ``replay.sample(256)`` + ``agent.train_ops(...)`` incl. priority write-back
and the amortised delayed actor step / target sync.  Prints ONE JSON line (rank 0).

* value   : updates/s with everything resident on the device (burst API ``train_ops_n``), timed with
            CUDA events, max over ranks; whole-job aggregate over N GPUs (independent agents per GPU:
            the per-agent update does not shard -- "replicas", weak scaling).
* e2e     : the same metric through the reference-facing Python API with HOST data in the loop: every
            step appends one transition from pinned host memory (H2D), samples, runs ``train_ops`` and
            reads the info floats back (D2H + sync), as ``rl/runner/run.py:160-171`` drives it.
* roofline: the dominant kernel (grouped FFMA GEMM) and, separately, the HBM-bound LAP tile-sum kernel.
* cpu_baseline / --impl reference: the oracle port of the reference's CPU update loop (the reference is
            pure Python and cannot travel to the GPU box), all host threads, bounded sample.
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
    # name: (algo, env, prioritized, seeds stacked per GPU, buffer rows)
    "td7_halfcheetah_1M": ("TD7", "HalfCheetah-v4", True, 1, 1_000_000),
    "td7_humanoid_1M": ("TD7", "Humanoid-v4", True, 1, 1_000_000),
    "td7_ant_1M": ("TD7", "Ant-v4", True, 1, 1_000_000),
    "td3_hopper_1M": ("TD3", "Hopper-v4", False, 1, 1_000_000),
    "sac_walker2d_1M": ("SAC", "Walker2d-v4", False, 1, 1_000_000),
    "td7_ant_pop256": ("TD7", "Ant-v4", True, 256, 1_000_000),
    "td7_humanoid_pop256": ("TD7", "Humanoid-v4", True, 256, 200_000),
}
DEFAULT_WORKLOAD = "td7_halfcheetah_1M"  # BASELINE.json configs[1]
ENV_DIMS = {"Hopper-v4": (11, 3), "HalfCheetah-v4": (17, 6), "Walker2d-v4": (17, 6), "Ant-v4": (27, 8), "Humanoid-v4": (376, 17)}
BATCH = 256


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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_updates_per_s(workload: str, budget_s: float, steps: int | None = None, warmup: int = 5):
    """The reference's CPU update loop (oracle port: same torch ops in the same order, pinned bit-exact to
    the reference by tests/test_oracle_golden.py) on this host, all threads."""
    import torch

    from oracle import agents as oagents
    from oracle import replay as oreplay

    algo, env, prio, _, rows = WORKLOADS[workload]
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    agent = {"TD7": oagents.OracleTD7, "TD3": oagents.OracleTD3, "SAC": oagents.OracleSAC}[algo](env, use_lap=prio)
    buf = (oreplay.OracleLAP if prio else oreplay.OracleUniform)(rows, env)
    buf.fill_synthetic(rows, seed=0)
    for _ in range(warmup):
        agent.train_ops(buf.sample(BATCH), replay_buffer=buf)
    n, t0 = 0, time.perf_counter()
    while True:
        agent.train_ops(buf.sample(BATCH), replay_buffer=buf)
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= budget_s):
            break
        if el > 180:
            break
    return n / el, n, el, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # A reference-arm step is a bounded sample of the workload: `per_step` consecutive updates, sized so that K steps hold
    # at least ~600 updates (8-10 s of CPU work; a handful of 14 ms updates would time thread wake-ups, not the loop) and
    # the whole run stays under a few minutes.  The default run (no --steps) is a 20 s sample.
    warm = max(args.warmup, 10)
    per_step = max(1, -(-600 // args.steps)) if args.steps_given else 1
    ups, n, el, threads = cpu_updates_per_s(args.workload, budget_s=20.0, steps=args.steps * per_step if args.steps_given else None, warmup=warm)
    algo, env, prio, seeds, rows = WORKLOADS[args.workload]
    k_steps = args.steps if args.steps_given else n
    line = {
        "impl": "reference", "metric": "gradient updates/sec (batch 256) per agent", "value": ups, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": k_steps, "updates_per_step": n / k_steps, "warmup": warm, "ms_per_step": 1e3 * el / k_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "algo": algo, "env": env, "batch": BATCH, "buffer_rows": rows, "replay": "LAP" if prio else "uniform"},
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": threads, "kind": "port",
                         "sample": f"{n} updates of the oracle port (torch {torch.__version__} CPU, {threads} threads, os.cpu_count={os.cpu_count()}) in {el:.1f}s"},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    def stop(self, timed=None, probe=None):
        """Median SM clock and the throttle reasons seen.  `timed` = (t0, t1) wall-clock bounds of the timed region: its
        samples are the ones reported; when the region is shorter than the 100 ms sampling period and holds none, the
        samples of the whole loaded window (warm-up + timed + end-to-end loop) are reported instead, and `window` says so."""
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
                ts = -1.0  # unparsed time stamp: the sample still counts for the whole window
            try:
                rows.append((ts, float(parts[1]), float(parts[2]), [nm for nm, v in zip(names, parts[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        window = "warm-up + timed region + end-to-end loop"
        if timed is not None:
            inside = [r for r in rows if timed[0] <= r[0] <= timed[1]]
            if inside:
                rows, window = inside, "timed region"
            elif probe is not None:
                inside = [r for r in rows if probe[0] <= r[0] <= probe[1]]
                if inside:
                    rows, window = inside, "untimed repeat of the same update loop for 0.6 s (the timed region is shorter than nvidia-smi's 100 ms period)"
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


def gemm_roofline(lib, torch, S, A, groups=1):
    """Time the largest grouped launch of the TD7 critic chain (csrc/agent_td7.cuh table C[7]: five
    independent layers, K up to 768) standalone with CUDA events."""
    from sac_td3_td7_b200 import _lib

    B, H, Z = BATCH, 256, 256
    dev = torch.device("cuda")
    mk = lambda *s: torch.randn(*s, device=dev) * 0.05  # noqa: E731
    probs, flops, keep = [], 0.0, []

    def add(xs, n_out):
        nonlocal flops
        K = sum(x.shape[1] for x in xs)
        ld = (K + 3) // 4 * 4
        W, b, y = mk(n_out, ld), mk(n_out), torch.empty(B, n_out, device=dev)
        keep.extend([W, b, y, *xs])
        p = _lib.GemmProblem()
        p.I, p.J, p.C, p.c_ld, p.act, p.bias = B, n_out, y.data_ptr(), n_out, 2, b.data_ptr()
        off = 0
        for x in xs:
            s = p.seg[p.nseg]
            s.A, s.a_ld, s.B, s.b_ld, s.len = x.data_ptr(), x.shape[1], W.data_ptr() + 4 * off, ld, x.shape[1]
            p.nseg += 1
            off += x.shape[1]
        probs.append(p)
        flops += 2.0 * B * n_out * K

    zs, a, s2, qn, zsa = mk(B, Z), mk(B, A), mk(B, S), mk(B, H), mk(B, Z)
    add([zs, a], H)            # fixed_encoder_target.zsa1
    add([s2, a], H)            # target_q1.q01
    add([s2, a], H)            # target_q2.q01
    add([qn, zsa, zs], H)      # q1.q1
    add([qn, zsa, zs], H)      # q2.q1
    arr = (_lib.GemmProblem * len(probs))(*probs)
    ms = C.c_float()
    _lib.check(lib.b200rl_op_gemm_timed(0, arr, len(probs), groups, int(os.environ.get("B200_GEMM_WARMUP", "20")), int(os.environ.get("B200_GEMM_ITERS", "200")), C.byref(ms), torch.cuda.current_stream().cuda_stream))
    return flops, ms.value


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sac_td3_td7_b200 import _lib, build
    from sac_td3_td7_b200.rl.agent import SAC, TD3, TD7
    from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the update path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    lib = _lib.lib()

    algo, env, prio, seeds, rows = WORKLOADS[args.workload]
    S, A = ENV_DIMS[env]
    K, W = args.steps, max(args.warmup, 3)
    torch.manual_seed(1000 + rank)
    cls = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo]
    agent = cls(env, use_lap=prio, n_seeds=seeds) if algo != "SAC" else cls(env, n_seeds=seeds)
    buf = (LAPReplayMemory if prio else SimpleReplayMemory)(rows, env, n_seeds=seeds, share_rows=seeds > 1)
    buf.fill_synthetic(rows, seed=rank)
    stream = torch.cuda.current_stream()

    # ---- device-resident throughput: K updates through the burst API ----
    clocks = ClockSampler(local) if rank == 0 else None
    agent.train_ops_n(W, buf)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.time()
    e0.record(stream)
    infos = agent.train_ops_n(K, buf)
    e1.record(stream)
    torch.cuda.synchronize()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    last = {k: (None if v is None else float(v)) for k, v in infos[-1].items()}
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms_max = float(t.item())
    value = world * seeds * K / (ms_max / 1e3)

    # ---- end to end through the drop-in API with host data in the loop ----
    e2e_steps = min(K, 2000)
    g = np.random.default_rng(5)
    pinned = torch.empty((64, buf.row_stride), dtype=torch.float32).pin_memory()
    pinned.copy_(torch.from_numpy(buf.pack(g.standard_normal((64, S), dtype=np.float32), g.random((64, A), dtype=np.float32) * 2 - 1,
                                           g.standard_normal((64, 1), dtype=np.float32), g.standard_normal((64, S), dtype=np.float32),
                                           np.ones((64, 1), np.float32))))
    pinned_np = pinned.numpy()

    def e2e_step(i):
        for s in range(seeds if not buf.share_rows else 1):
            buf.append_rows(pinned_np[i % 64:i % 64 + 1], s)            # H2D of one transition from pinned memory
        batch = buf.sample(BATCH)
        info = agent.train_ops(batch, replay_buffer=buf)
        return float(info["train/q_fn"])                                   # D2H of the info row + sync

    e2e = None
    if seeds == 1:
        for i in range(W):
            e2e_step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_steps / float(t.item()), "unit": "updates/s", "h2d_bytes_per_step": buf.row_stride * 4,
               "d2h_bytes_per_step": 8 * len(agent.INFO_KEYS), "steps": e2e_steps,
               "note": "per step: append 1 transition from pinned host memory, sample(256), train_ops, float(info['train/q_fn']) -- the engine posts "
                       "every loss of the step to a pinned host mailbox the moment its kernel has run (8 B per value), so the read returns while "
                       "the step's backward pass is still executing and the next step's host calls overlap it; full sync at the end of the loop"}
    probe = None
    if wall1 - wall0 < 0.35:  # too short for nvidia-smi's 100 ms period: sample the same loop, untimed, for 0.6 s more
        p0 = time.time()
        while time.time() - p0 < 0.6:
            agent.train_ops_n(max(1, min(K, 200)), buf)
            torch.cuda.synchronize()
        probe = (p0, time.time())
    clk = clocks.stop(timed=(wall0, wall1), probe=probe) if clocks else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (rank 0, standalone event-timed launches) ----
    hbm_peak, bf16_burst, bf16_sus, peak_src = measured_peaks()
    g_flops, g_ms = gemm_roofline(lib, torch, S, A)
    ach_tf = g_flops / (g_ms * 1e-3) / 1e12
    roofline = {"kernel": "gemm_kernel<fwd> (grouped FFMA GEMM, 5 independent layers of the TD7 critic chain, 320 CTAs)", "bound": "tensor",
                "achieved": ach_tf, "peak": bf16_burst, "unit": "TFLOP/s", "frac": ach_tf / bf16_burst,
                "traffic": 2.7392e6,  # dram__bytes_read + write of this launch, profiles/gemm_ffma_ncu_full_r01.summary.txt (operands are L2-resident)
                "peak_source": peak_src + ", dense bf16 cuBLAS burst",
                "note": "fp32 FFMA by design (1e-5 fp32 parity rules out single-pass TF32/BF16); the non-tensor fp32 peak of a B200 is ~74.5 TFLOP/s",
                "fp32_ffma_peak_tflops": 74.5, "frac_of_fp32_ffma_peak": ach_tf / 74.5, "launch_ms": g_ms, "flops_per_launch": g_flops}
    roof2 = None
    if prio:
        ms2 = (C.c_float * 2)()
        _lib.check(lib.b200rl_replay_time_sample(buf._h, BATCH, 10, 200, ms2, stream.cuda_stream))
        bytes_scan = 4.0 * rows * seeds
        row_bytes = 4.0 * (2 * S + A + 2)
        nsuper = -(-rows // (256 * 64))
        # per draw: the super sums, 64 tile sums, 256 priorities, the packed row in and the five batch tensors out, the index
        bytes_walk = seeds * BATCH * (nsuper * 8 + 64 * 8 + 256 * 4 + row_bytes * 2 + 8 + 4)
        # dram__bytes_read + write of one launch from the committed ncu --set full captures (profiles/lap_sample_*_ncu_full_r01.summary.txt)
        ncu_traffic = {"td7_halfcheetah_1M": 0.379e6, "td7_humanoid_pop256": 422.8e6, "td7_ant_pop256": 95.8e6}.get(args.workload)
        roof2 = {"kernel": "lap_sample_kernel (per draw: walk of the 3-level fp64 sum tree + row gather; the only replay kernel on the per-step path)",
                 "bound": "hbm", "achieved": bytes_walk / (ms2[1] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": bytes_walk / (ms2[1] * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic, "launch_ms": ms2[1], "bytes_per_launch": bytes_walk, "peak_source": peak_src,
                 "note": "latency-bound by construction for one agent (256 warps, three dependent L2 round trips each); the tree itself is repaired in place by the "
                         "kernels that write priorities, so no per-step pass over the 4 MB priority array remains.  Stacked seeds (256 x 256 draws per launch): "
                         "HBM-bound on Humanoid rows (ncu: 423 MB of DRAM traffic in 87 us = 59 % of ncu's DRAM peak, 74 % of the measured copy peak), "
                         "issue-bound on Ant rows (663 instructions per draw, issue slots 74 % busy)",
                 "full_rebuild_kernel": {"kernel": "lap_tile_sums_kernel (bulk loads / restored checkpoints only)", "launch_ms": ms2[0], "bytes_per_launch": bytes_scan,
                                         "achieved_gbs": bytes_scan / (ms2[0] * 1e-3) / 1e9}}

    roof3 = None
    if seeds >= 8:  # stacked seeds: the tcgen05 kernels (HBM- and tensor-side figures)
        roof3 = {}
        for mode, name in ((0, "forward"), (2, "wgrad_adam")):
            fl, by, ms3 = stacked_roofline(lib, torch, seeds, mode)
            roof3[name] = {"kernel": "gemm_tc_ws_kernel (persistent tcgen05 TF32, TMA-fed), 2 layers x %d seeds" % seeds, "launch_ms": ms3,
                           "hbm": {"achieved": by / (ms3 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": by / (ms3 * 1e-3) / 1e9 / hbm_peak},
                           "tensor": {"achieved": fl / (ms3 * 1e-3) / 1e12, "peak": bf16_sus, "unit": "TFLOP/s", "frac": fl / (ms3 * 1e-3) / 1e12 / bf16_sus,
                                      "note": "TF32 issue rate is half of bf16; the stacked update is HBM-bound at B=256 per seed"},
                           "bytes_per_launch": by, "flops_per_launch": fl, "peak_source": peak_src}

    # ---- CPU baseline (oracle port) on this host, bounded ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ups, n, el, threads = cpu_updates_per_s(args.workload if seeds == 1 else "td7_ant_1M", budget_s=15.0)
        cpu = {"value": ups, "unit": "updates/s", "cores": threads, "kind": "port",
               "sample": f"{n} updates of the oracle port of the reference CPU loop ({threads} torch threads, os.cpu_count={os.cpu_count()}) in {el:.1f}s, same workload"}

    lp0, lp1 = lib.b200rl_agent_launches_per_update(agent._h, 0), lib.b200rl_agent_launches_per_update(agent._h, 1)
    per_update = (lp0 + lp1) / 2.0 if algo != "SAC" else float(lp1)
    gflop = algorithmic_gflop(algo, S, A)
    line = {
        "metric": "gradient updates/sec (batch 256) per agent" if seeds == 1 else "agent-updates/sec (batch 256)",
        "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "algo": algo, "env": env, "obs_dim": S, "act_dim": A, "batch": BATCH, "buffer_rows": rows,
                   "replay": "LAP" if prio else "uniform", "seeds_per_gpu": seeds, "parallelism": f"replicas x{world} (independent agents, no data-path collective)",
                   "l2": "replay rows exceed L2 (1M rows); weights, Adam moments and the 4 MB priority array are L2-resident by design -- they are the loop state"},
        "e2e": e2e, "gpu_launches": int(K * (per_update + 1)), "launches_per_update": {"critic_only": lp0, "with_actor": lp1},
        "clocks": clk, "roofline": roofline, "roofline_replay": roof2, "roofline_stacked": roof3, "cpu_baseline": cpu,
        "algorithmic_gflop_per_update": gflop, "update_tflops_per_gpu": gflop * value / 1e3 / world,
        "last_info": last,
    }
    if cpu:
        line["speedup_vs_cpu_port"] = {"device_resident": value / cpu["value"], "e2e": (e2e["value"] / cpu["value"]) if e2e else None}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stacked_roofline(lib, torch, G, mode):
    """Graph-timed launch of the tcgen05 grouped GEMM on `G` stacked seeds: two 256x256x256 layers (mode 0: forward +
    bias + ELU; mode 2: weight gradient + fused Adam).  Returns (flops, algorithmic bytes, ms per launch)."""
    from sac_td3_td7_b200 import _lib

    os.environ.setdefault("B200RL_TC", "1")
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 2000
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
