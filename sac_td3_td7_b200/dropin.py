"""Option B of INTEGRATION.md: make an UNMODIFIED checkout of the reference train through this package.

    import sac_td3_td7_b200.dropin as dropin
    dropin.install()                      # before anything imports the reference's ``rl.agent`` / ``rl.replay_memory`` / ``rl.nn``
    from rl.agent import TD7              # -> the B200 TD7;  rl.runner / rl.rollout / rl.cli stay the reference's own

The reference has no plugin registry: ``run_td7`` & co. construct ``rl.agent.TD7`` and
``rl.replay_memory.LAPReplayMemory`` by name, and the runners import ``rl.agent.abc.Agent`` /
``rl.replay_memory.base.REPLAYMEMORY`` for annotations.  ``install()`` registers this package's ``rl.agent``,
``rl.replay_memory`` and ``rl.nn`` -- and every submodule of them -- under those names in ``sys.modules``; the rest of the
reference package (runners, rollout, CLI, utils) is imported from the reference as usual.
"""
from __future__ import annotations

import importlib
import pkgutil
import sys


def install(names=("replay_memory", "nn", "agent")) -> list[str]:
    """Alias ``sac_td3_td7_b200.rl.<name>`` (and its submodules) as ``rl.<name>``.  Returns the registered module names."""
    from . import rl as ours

    done = []
    for name in names:
        pkg = importlib.import_module(f"{ours.__name__}.{name}")
        sys.modules[f"rl.{name}"] = pkg
        done.append(f"rl.{name}")
        for info in pkgutil.iter_modules(pkg.__path__):
            sub = importlib.import_module(f"{pkg.__name__}.{info.name}")
            sys.modules[f"rl.{name}.{info.name}"] = sub
            done.append(f"rl.{name}.{info.name}")
    return done
