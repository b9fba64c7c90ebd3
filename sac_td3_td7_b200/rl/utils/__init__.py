from .annotation import BATCH, EPS  # noqa: F401
from ...envs import get_action_bias_scale, get_state_action_dims  # noqa: F401
