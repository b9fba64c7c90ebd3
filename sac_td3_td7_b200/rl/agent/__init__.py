from .abc import AGENT, Agent, LazyScalar
from .entry import run_sac, run_td3, run_td7
from .sac import SAC
from .td3 import TD3
from .td7 import TD7

__all__ = ["run_sac", "run_td3", "run_td7", "SAC", "TD3", "TD7", "Agent", "AGENT", "LazyScalar"]
