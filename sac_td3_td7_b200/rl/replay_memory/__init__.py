from .base import REPLAYMEMORY, BaseReplayMemory
from .device import LAPReplayMemory, SimpleReplayMemory

__all__ = ["SimpleReplayMemory", "LAPReplayMemory", "BaseReplayMemory", "REPLAYMEMORY"]
