"""The reference's OWN runner code driving the drop-in (INTEGRATION.md Option B): ``rl/rollout.py`` fills the buffer through
``append``, ``rl/runner/run.py:87-96`` ``run_train_ops`` trains, ``run.py:53-84`` ``log_train_infos`` formats the infos --
all imported from the staged unmodified copy ``baseline/_ref`` (written by ``baseline/fetch_ref.py`` in the build container;
it travels with the snapshot), with only ``rl.agent`` / ``rl.replay_memory`` / ``rl.nn`` aliased to this package."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

SCRIPT = r"""
import io, logging, sys
import numpy as np
ROOT, REF, ALGO = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
from oracle import gym_stub                      # fake gymnasium: MuJoCo is not installed (test infrastructure)
gym_stub.install()
import sac_td3_td7_b200.dropin as dropin
dropin.install()                                 # Option B: before the reference is imported
sys.path.insert(0, REF)
from rl.runner.run import run_train_ops, log_train_infos          # the reference's own functions
from rl.rollout import Rollout                                     # the reference's own class
from rl import agent as rl_agent
assert rl_agent.__name__.startswith("sac_td3_td7_b200") and "baseline/_ref" in sys.modules["rl.runner.run"].__file__.replace("\\", "/")
from rl.agent import TD7, TD3, SAC
from rl.replay_memory import LAPReplayMemory, SimpleReplayMemory

class ToyEnv:                                    # gymnasium API; HalfCheetah shapes
    def __init__(self):
        self.g = np.random.default_rng(0)
        self.action_space = gym_stub._Box(6, 1.0)
        self.t = 0
    def reset(self, **k):
        self.t = 0
        return self.g.standard_normal(17), {}
    def step(self, a):
        assert np.asarray(a).shape == (6,)
        self.t += 1
        return self.g.standard_normal(17), float(-np.square(a).sum()), False, self.t % 50 == 0, {}

env_id = "HalfCheetah-v4"
if ALGO == "TD7":
    agent, buf = TD7(env_id, use_lap=True), LAPReplayMemory(10000, env_id)
elif ALGO == "TD3":
    agent, buf = TD3(env_id), SimpleReplayMemory(10000, env_id)
else:
    agent, buf = SAC(env_id), SimpleReplayMemory(10000, env_id)
rollout = Rollout(ToyEnv(), buf)
for _ in range(200):
    rollout.sample()                             # RandomSampler (rl/sampler.py) -> append([...]) (rl/rollout.py:45)
rollout.set_sampler(agent)                       # then the agent acts (agent.sample(obs))
for _ in range(312):
    rollout.sample()
assert len(buf) == 512 and buf.size == 512
infos = run_train_ops(rollout, agent, 256, n_ops=50)
assert len(infos) == 50
stream = io.StringIO()
logger = logging.getLogger("t"); logger.setLevel(logging.INFO); logger.addHandler(logging.StreamHandler(stream))
log_train_infos(50, logger, infos, {"perf/mean": 1.0}, {"rollout/n_episode": 10.0}, True)
lines = stream.getvalue().strip().splitlines()
assert lines[0].startswith("iteration,") and lines[1].startswith("50,"), lines
vals = [float(x) for x in lines[1].split(",")[1:]]
assert len(vals) == len(lines[0].split(",")) - 1 and all(np.isfinite(vals)), lines
print("OK", ALGO, lines[0], lines[1])
"""


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["TD7", "TD3", "SAC"])
def test_reference_runner_code_trains_the_dropin(algo):
    if not os.path.isdir(os.path.join(REF, "rl")):
        pytest.skip("baseline/_ref is not staged (run `python baseline/fetch_ref.py` where /root/reference exists)")
    res = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF, algo], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "OK " + algo in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
