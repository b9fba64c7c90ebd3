from .abc import ACTOR, CRITIC, ENCODER
from .modules import AvgL1Norm, MLPActor, MLPCritic, SALEActor, SALECritic, SALEEncoder, make_mlp
from .utils import annotate_make_nn, annotate_make_nn_td7, calculate_grad_norm, clip_grad_norm

__all__ = ["ACTOR", "CRITIC", "ENCODER", "MLPActor", "MLPCritic", "SALEActor", "SALECritic", "SALEEncoder", "AvgL1Norm", "make_mlp",
           "calculate_grad_norm", "clip_grad_norm", "annotate_make_nn", "annotate_make_nn_td7"]
