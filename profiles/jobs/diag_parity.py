"""Diagnostic (not a test): per-step, per-tensor gradient errors of the teacher-forced parity cases on one product loop.
    python profiles/jobs/diag_parity.py mma|ffma case [case ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import devlib, golden_util as gu
import test_agent_parity as T

dev = devlib.Dev("cuda")
dev.lib.b200rl_set_mma(1 if sys.argv[1] == "mma" else 0)
torch.set_num_threads(1)
for name in sys.argv[2:]:
    algo, g, oracle, obuf, h, r, S, A, steps, prioritized = T.setup_case(dev, name)
    for t in range(steps):
        T.push_state(dev, h, algo, oracle)
        if prioritized:
            dev.replay_write_priority(r, obuf.priority[: obuf.size].numpy(), max_priority=float(obuf.max_priority))
        u, noise = g[f"s{t}/u"], gu.step_noise(g, t, algo)
        obatch = obuf.sample(gu.BATCH, u=torch.tensor(u))
        oinfo = oracle.train_ops(obatch, replay_buffer=obuf, noise=noise)
        idx, batch = T.c_step(dev, h, r, S, A, u, noise, algo)
        info = dev.read_info(h, t + 1)
        print(name, "step", t, "info", [float(x) for x in info], {k: v for k, v in oinfo.items()})
        print("   q_target err", T.nerr(dev.debug(h, 0, gu.BATCH), oracle.last["q_target"].numpy().reshape(-1)), "q1 err", T.nerr(dev.debug(h, 2, gu.BATCH), oracle.last["q1"].numpy().reshape(-1)))
        for nid, (nm, layers) in enumerate(T.net_table(algo)):
            key = "grad_" + nm
            if nm not in T.TRAINED or key not in oracle.last:
                continue
            got = dev.read_net(h, nid, which=3)
            for (w, b), ln in zip(got, layers):
                for arr, suffix in ((w, ".weight"), (b, ".bias")):
                    ref = oracle.last[key][ln + suffix]
                    if ref is None:
                        continue
                    ref = ref.numpy()
                    d = np.abs(arr.astype(np.float64) - ref).reshape(arr.shape[0], -1).max(1) / max(np.abs(ref).max(), 1e-30)
                    bad = int((d > 1e-5).sum())
                    print(f"   grad {nm}.{ln}{suffix}: max {d.max():.2e} median {np.median(d):.2e} rows>1e-5: {bad}/{d.size}" + (f" worst rows {np.argsort(d)[-3:]}" if bad else ""))
        if algo == "TD7":
            ref4 = np.array([oracle.value_max, oracle.value_min, oracle.value_target_max, oracle.value_target_min])
            dev.check(dev.lib.b200rl_td7_set_value_range(h, 0, ref4.astype(np.float32).ctypes.data, None))
