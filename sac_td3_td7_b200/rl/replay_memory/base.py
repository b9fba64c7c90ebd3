"""ABC of the replay memories (mirrors ``rl/replay_memory/base.py:9-26``)."""
from abc import ABC
from typing import Any, TypeVar

REPLAYMEMORY = TypeVar("REPLAYMEMORY", bound="BaseReplayMemory")


class BaseReplayMemory(ABC):
    def __init__(self, replay_buffer_size: int, **kwargs) -> None:
        self.replay_buffer_size = replay_buffer_size

    def sample(self, batch_size: int, use_torch: bool = True):
        raise NotImplementedError("Sample method should be implemented.")

    def append(self, transition: list[Any]) -> None:
        raise NotImplementedError("Append method should bed implemented.")

    def __len__(self) -> int:
        raise NotImplementedError("__len__ should be implemented.")
