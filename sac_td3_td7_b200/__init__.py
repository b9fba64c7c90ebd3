"""B200-native (sm_100a) off-policy update path for SAC / TD3 / TD7.

``sac_td3_td7_b200.rl`` mirrors the Python API of the reference's ``rl`` package for the hot
path only -- ``rl.replay_memory`` (append / sample / update_priority / reset_max_priority),
``rl.nn`` (SALE and MLP modules) and ``rl.agent`` (TD7 / TD3 / SAC ``train_ops``) -- so the
reference's ``rl/runner`` and ``rl/rollout`` can drive it unchanged.
All device work goes through the C ABI of ``libb200rl.so`` (``include/b200rl.h``); there is no
CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["rl", "_lib", "build", "population"]
