"""ACTOR / CRITIC / ENCODER bases (method names the agents call; ``rl/nn/abc.py:7-46``)."""
from abc import ABC

import torch
from torch import nn


class ACTOR(ABC, nn.Module):
    def inference_mean(self, *args, **kwargs) -> torch.Tensor:
        pass

    def inference_mean_logvar(self, *args, **kwargs):
        pass

    def inference_policy(self, *args, **kwargs) -> torch.Tensor:
        pass


class CRITIC(ABC, nn.Module):
    def estimate_state_value(self, *args, **kwargs) -> torch.Tensor:
        pass

    def estimate_q_value(self, *args, **kwargs) -> torch.Tensor:
        pass


class ENCODER(ABC, nn.Module):
    def encode_state(self, *args, **kwargs) -> torch.Tensor:
        pass

    def encode_state_action(self, *args, **kwargs) -> torch.Tensor:
        pass
