"""train.log / eval.log writers: byte-identical to what the reference's runner would log (rl/runner/run.py:47-84,
157,180-181), restated here line by line for the check (host logic only; no GPU)."""
import numpy as np

from sac_td3_td7_b200.runlog import PopulationLogger, RunLogger, mean_with_dirty_eles, summarize_returns


def reference_train_lines(episodes):
    """run.py:53-84 restated: returns the lines the reference's train logger would emit."""
    lines, started = [], False
    for iteration, train_infos, test_info, rollout_info in episodes:
        keys = list(train_infos[0].keys())
        info = {}
        for k in keys:
            eles = [x[k] for x in train_infos if x[k] is not None]
            info[k] = sum(eles) / len(eles) if len(eles) != 0 else -1e6
        if not started:
            lines.append(",".join(["iteration"] + sorted(list(info.keys()) + list(rollout_info.keys()) + list(test_info.keys()))))
            started = True
        info.update(rollout_info)
        info.update(test_info)
        info = {k: v for k, v in sorted(info.items())}
        lines.append(f"{iteration}," + ",".join(f"{v:.4f}" for v in info.values()))
    return lines


def test_train_and_eval_logs_match_reference_schema(tmp_path):
    rng = np.random.default_rng(0)
    episodes = []
    it = 0
    for ep in range(5):
        n = int(rng.integers(1, 40))
        infos = [{"train/encoder": float(rng.random()), "train/q_fn": float(rng.random() * 3),
                  "train/policy": (float(rng.standard_normal()) if (it + k) % 2 == 0 else None)} for k in range(n)]
        it += n
        test_info = {"perf/mean": float(rng.random() * 1000), "perf/min": 1.5, "perf/max": 2000.25}
        rollout_info = {"rollout/return": float(rng.random() * 500), "rollout/episode_length": float(n)}
        episodes.append((it, infos, test_info, rollout_info))
    # an episode whose updates never ran the actor: the reference logs -1e6 for that key
    episodes.append((it + 1, [{"train/encoder": 0.5, "train/q_fn": 1.0, "train/policy": None}], episodes[-1][2], episodes[-1][3]))
    lg = RunLogger(tmp_path / "run")
    for e in episodes:
        lg.log_train(*e)
    evals = [(1000, {"perf/mean": 12.3456, "perf/min": -1.0, "perf/max": 99.9996}), (2000, {"perf/mean": 3.0, "perf/min": 2.0, "perf/max": 4.0})]
    for t, info in evals:
        lg.log_eval(t, info)
    lg.close()
    got = (tmp_path / "run" / "train.log").read_text().splitlines()
    assert got == reference_train_lines(episodes)
    assert "-1000000.0000" in got[-1]
    got_eval = (tmp_path / "run" / "eval.log").read_text().splitlines()
    assert got_eval == ["timestep,perf/mean,perf/min,perf/max", "1000,12.346,-1.000,100.000", "2000,3.000,2.000,4.000"]
    assert mean_with_dirty_eles([None, None]) == -1e6 and mean_with_dirty_eles([1.0, None, 3.0]) == 2.0


class _FakeAgent:
    def info_dicts(self, first_step, count, seed=0):
        return [{"train/q_fn": float(seed + first_step + i), "train/policy": None if (first_step + i) % 2 else float(-seed)} for i in range(count)]


def test_population_logger_writes_one_pair_per_owned_seed(tmp_path):
    pl = PopulationLogger(tmp_path, first_seed=4, count=3)
    ti = [summarize_returns([1.0 * g, 2.0 * g + 1]) for g in range(3)]
    ri = [{"rollout/return": 10.0 + g} for g in range(3)]
    pl.log_train(7, _FakeAgent(), first_step=5, n_steps=4, test_infos=ti, rollout_infos=ri)
    pl.log_eval(5000, [[1.0, 3.0], [2.0, 2.0], [0.0, 10.0]])
    pl.close()
    for g in range(3):
        d = tmp_path / f"seed_{4 + g}"
        tr = (d / "train.log").read_text().splitlines()
        assert tr[0] == "iteration,perf/max,perf/mean,perf/min,rollout/return,train/policy,train/q_fn"
        q = np.mean([g + 5 + i for i in range(4)])
        assert tr[1] == f"7,{ti[g]['perf/max']:.4f},{ti[g]['perf/mean']:.4f},{ti[g]['perf/min']:.4f},{10.0 + g:.4f},{float(-g):.4f},{q:.4f}"
        ev = (d / "eval.log").read_text().splitlines()
        assert ev[0] == "timestep,perf/mean,perf/min,perf/max" and ev[1].startswith("5000,")
    assert summarize_returns([1.0, 3.0]) == {"perf/mean": 2.0, "perf/min": 1.0, "perf/max": 3.0}
