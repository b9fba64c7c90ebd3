"""``rl/nn/utils.py`` helpers that touch the hot path."""
from functools import wraps

import torch
from torch import nn


def calculate_grad_norm(neural_network: nn.Module) -> float:
    """rl/nn/utils.py:13-19 -- sum over tensors of ||grad||_2 (TD3 logs it as norm/policy; the
    update engine computes the same quantity on the device, see csrc/ops.cuh grad_norm_sum_kernel)."""
    total = 0.0
    for p in neural_network.parameters():
        if p.grad is not None:
            total += torch.norm(p.grad.data, p=2)
    return total.item() if isinstance(total, torch.Tensor) else total


def clip_grad_norm(neural_network: nn.Module, max_norm: float = 1.0) -> None:
    if max_norm == float("inf"):
        return
    torch.nn.utils.clip_grad_norm_(neural_network.parameters(), max_norm)


def annotate_make_nn(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        return func(*args, **kwargs)

    return wrapper


annotate_make_nn_td7 = annotate_make_nn
