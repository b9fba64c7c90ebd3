"""Drop-in mirror of the reference's ``rl`` package for the update hot path."""
from . import agent, nn, replay_memory, utils  # noqa: F401

__all__ = ["agent", "nn", "replay_memory", "utils"]
