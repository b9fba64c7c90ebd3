"""Constants / contracts of ``rl/utils/annotation.py`` used on the hot path."""
from typing import TypedDict

import torch

EPS = 1e-6  # rl/utils/annotation.py:12 (SAC log-prob)


class BATCH(TypedDict):
    """rl/utils/annotation.py:23-30"""

    state: torch.Tensor
    next_state: torch.Tensor
    action: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
