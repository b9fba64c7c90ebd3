"""``train.log`` / ``eval.log`` writers with the reference's CSV schema, for one run or a seed population.

The reference logs through ``logging`` handlers that emit bare messages (rl/utils/miscellaneous.py
``setup_logger``), so the files are plain CSV:

* ``train.log`` (rl/runner/run.py:53-84 ``log_train_infos``): header ``iteration,<keys>`` once, keys = the
  union of train / rollout / test info keys SORTED; one row per episode ``<iteration>,<v:.4f>,...`` with the
  values in the same sorted order; a train key is the mean over the updates of the episode that produced
  a value (``None`` entries dropped, ``-1e6`` when none did -- run.py:47-51).
* ``eval.log`` (run.py:157,180-181): header ``timestep,perf/mean,perf/min,perf/max``; rows
  ``<timestep>,<v:.3f>,...`` in the order of the test-info dict.

A population writes one pair of files per seed under ``<base_dir>/seed_<k>/``; every rank writes the seeds it
owns (sac_td3_td7_b200/population.py::shard_seeds), so no gather is needed for logging.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, Mapping, Sequence

import numpy as np


def mean_with_dirty_eles(eles: Iterable) -> float:
    """run.py:47-51"""
    vals = [float(e) for e in eles if e is not None]
    return sum(vals) / len(vals) if vals else -1e6


class RunLogger:
    """The two log files of one (seed of a) run."""

    def __init__(self, base_dir, resume: bool = False):
        self.base_dir = Path(base_dir)
        self.base_dir.mkdir(parents=True, exist_ok=True)
        mode = "a" if resume else "w"
        self._train = open(self.base_dir / "train.log", mode, buffering=1)
        self._eval = open(self.base_dir / "eval.log", mode, buffering=1)
        self._train_started = resume
        self._eval_started = resume

    def log_train(self, iteration: int, train_infos: Sequence[Mapping], test_info: Mapping, rollout_info: Mapping) -> None:
        info = {k: mean_with_dirty_eles(ti[k] for ti in train_infos) for k in train_infos[0].keys()}
        if not self._train_started:
            self._train.write(",".join(["iteration"] + sorted(list(info) + list(rollout_info) + list(test_info))) + "\n")
            self._train_started = True
        info.update(rollout_info)
        info.update(test_info)
        self._train.write(f"{iteration}," + ",".join(f"{float(v):.4f}" for _, v in sorted(info.items())) + "\n")

    def log_eval(self, timestep: int, test_info: Mapping) -> None:
        if not self._eval_started:
            self._eval.write(",".join(["timestep"] + list(test_info.keys())) + "\n")
            self._eval_started = True
        self._eval.write(f"{int(timestep)}," + ",".join(f"{float(v):.3f}" for v in test_info.values()) + "\n")

    def close(self) -> None:
        self._train.close()
        self._eval.close()


def summarize_returns(returns: Sequence[float]) -> dict:
    """run.py:36-45: the evaluation summary of one agent from its episode returns."""
    r = np.asarray(returns, np.float64)
    return {"perf/mean": float(r.mean()), "perf/min": float(r.min()), "perf/max": float(r.max())}


class PopulationLogger:
    """One RunLogger per seed owned by this rank."""

    def __init__(self, base_dir, first_seed: int, count: int, resume: bool = False):
        self.first, self.count = int(first_seed), int(count)
        self.loggers = [RunLogger(Path(base_dir) / f"seed_{first_seed + g}", resume) for g in range(count)]

    def log_train(self, iteration: int, agent, first_step: int, n_steps: int, test_infos: Sequence[Mapping], rollout_infos: Sequence[Mapping]) -> None:
        """Episode summary of every local seed from the engine's info ring (updates first_step .. first_step+n_steps-1)."""
        for g, lg in enumerate(self.loggers):
            lg.log_train(iteration, agent.info_dicts(first_step, n_steps, seed=g), test_infos[g], rollout_infos[g])

    def log_eval(self, timestep: int, returns_per_seed: Sequence[Sequence[float]]) -> None:
        for g, lg in enumerate(self.loggers):
            lg.log_eval(timestep, summarize_returns(returns_per_seed[g]))

    def close(self) -> None:
        for lg in self.loggers:
            lg.close()
