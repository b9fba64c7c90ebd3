from .abc import AGENT, Agent, LazyScalar
from .sac import SAC
from .td3 import TD3
from .td7 import TD7

__all__ = ["SAC", "TD3", "TD7", "Agent", "AGENT", "LazyScalar"]
