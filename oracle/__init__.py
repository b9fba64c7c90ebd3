"""CPU oracle for the off-policy update hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch-fp32 / numpy) *restatement* of the reference's
algorithm for the path named by ``BASELINE.json:north_star``:

* ``oracle.replay``  -- LAP / uniform ring replay (``rl/replay_memory/{lap,simple}.py``)
* ``oracle.nets``    -- SALE + MLP forward maps (``rl/nn/{sale,mlp}.py``)
* ``oracle.agents``  -- TD7 / TD3 / SAC ``train_ops`` (``rl/agent/{td7,td3,sac}.py``)
* ``oracle.envs``    -- the observation/action shapes of the five MuJoCo tasks

Parity status: **pinned against the reference itself**.  The reference ships no
tests or golden vectors, so ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` (under
``oracle.gym_stub``) in the build container and records its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds this oracle to those
vectors bit-for-bit.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  The product package ``sac_td3_td7_b200`` never imports it.
"""
