"""``run_td7`` / ``run_td3`` / ``run_sac`` -- the names ``rl/cli.py:7`` imports from ``rl.agent``.

Only meaningful under INTEGRATION.md Option B (``sac_td3_td7_b200.dropin.install()`` inside a checkout of the reference):
the environment, the runners, seeding, logging and the run directory are the REFERENCE's own code (``rl.utils.make_env``,
``rl.runner.run_rl`` / ``run_rl_w_ckpt``, ``rl.utils.miscellaneous.fix_seed``, ``rl.SAVE_DIR``), imported when called; this
module contributes what ``rl/agent/td7.py:338-399`` / ``td3.py`` / ``sac.py:301-350`` decide themselves -- which replay
class and which agent class to construct -- with this package's classes."""
from __future__ import annotations

from datetime import datetime


def _launch(algo: str, run_name: str, env_id: str, seed: int, prioritized: bool, replay_buffer_size: int, checkpoints: bool, agent_kwargs: dict, runner_kwargs: dict):
    from rl import SAVE_DIR  # the reference package (control plane: out of this repository's scope)
    from rl.runner import run_rl, run_rl_w_ckpt
    from rl.utils import make_env
    from rl.utils.miscellaneous import fix_seed

    from ..replay_memory import LAPReplayMemory, SimpleReplayMemory
    from . import SAC, TD3, TD7

    base_dir = SAVE_DIR / algo / f"{run_name}-{datetime.strftime(datetime.now(), '%Y-%m-%d-%H:%M:%S')}"
    base_dir.mkdir(exist_ok=True, parents=True)
    fix_seed(seed)
    env = make_env(env_id)
    env.reset(seed=seed)
    replay_buffer = (LAPReplayMemory if prioritized else SimpleReplayMemory)(replay_buffer_size, env_id)
    agent = {"TD7": TD7, "TD3": TD3, "SAC": SAC}[algo](env_id, **agent_kwargs)
    (run_rl_w_ckpt if checkpoints else run_rl)(env, agent, replay_buffer, base_dir, **runner_kwargs)


def run_td7(run_name: str, env_id: str, seed: int = 777, without_policy_checkpoint: bool = False, without_lap: bool = False,
            replay_buffer_size: int = 1_000_000, record_video: bool = False, show_progressbar: bool = True, **kwargs) -> None:
    """td7.py:338-399"""
    _launch("TD7", run_name, env_id, seed, not without_lap, replay_buffer_size, not without_policy_checkpoint, dict(use_lap=not without_lap, **kwargs),
            dict(record_video=record_video, show_progressbar=show_progressbar, **kwargs))


def run_td3(run_name: str, env_id: str, seed: int = 777, use_lap: bool = False, replay_buffer_size: int = 1_000_000, record_video: bool = False,
            show_progressbar: bool = True, **kwargs) -> None:
    """td3.py:248-300 (use_lap selects the buffer AND the agent's loss)"""
    _launch("TD3", run_name, env_id, seed, use_lap, replay_buffer_size, False, dict(use_lap=use_lap, **kwargs),
            dict(record_video=record_video, show_progressbar=show_progressbar, **kwargs))


def run_sac(run_name: str, env_id: str, use_lap: bool = False, replay_buffer_size: int = 1_000_000, record_video: bool = False, seed: int = 777, **kwargs) -> None:
    """sac.py:301-350 (use_lap only selects the buffer: SURVEY B.7)"""
    _launch("SAC", run_name, env_id, seed, use_lap, replay_buffer_size, False, dict(**kwargs), dict(record_video=record_video, **kwargs))
