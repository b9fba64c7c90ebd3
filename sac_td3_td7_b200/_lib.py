"""ctypes binding of ``libb200rl.so`` (the C ABI declared in ``include/b200rl.h``).

The library is built in-tree by :mod:`sac_td3_td7_b200.build` (nvcc, sm_100a).  There is no CPU
fallback: :func:`lib` raises if the shared object is missing, and every entry point that touches
the device fails with ``B200RL_ERR_NO_DEVICE`` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200RL_LIB selects another build of the same sources (profiles/: the -DB200RL_PROBE timeline build)
LIB_PATH = os.environ.get("B200RL_LIB") or os.path.join(_HERE, "libb200rl.so")

c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)
c_vp = C.c_void_p


class TD7Config(C.Structure):
    _fields_ = [
        ("state_dim", C.c_int), ("action_dim", C.c_int), ("hdim", C.c_int), ("zs_dim", C.c_int),
        ("batch", C.c_int), ("groups", C.c_int),
        ("discount", C.c_double), ("policy_lr", C.c_double), ("critic_lr", C.c_double),
        ("target_update_rate", C.c_int), ("policy_freq", C.c_int),
        ("target_policy_noise", C.c_double), ("noise_clip", C.c_double),
        ("use_lap", C.c_int), ("keep_grads", C.c_int),
    ]


class MlpConfig(C.Structure):
    _fields_ = [
        ("algo", C.c_int), ("state_dim", C.c_int), ("action_dim", C.c_int), ("hdim", C.c_int),
        ("batch", C.c_int), ("groups", C.c_int),
        ("discount", C.c_double), ("policy_lr", C.c_double), ("critic_lr", C.c_double), ("tau", C.c_double),
        ("policy_freq", C.c_int),
        ("target_policy_noise", C.c_double), ("noise_clip", C.c_double),
        ("min_log_std", C.c_double), ("max_log_std", C.c_double), ("tmp", C.c_double),
        ("use_lap", C.c_int), ("keep_grads", C.c_int),
    ]


class GemmSeg(C.Structure):
    _fields_ = [("A", c_vp), ("B", c_vp), ("a_gs", C.c_int64), ("b_gs", C.c_int64),
                ("a_ld", C.c_int), ("b_ld", C.c_int), ("len", C.c_int), ("b_ones", C.c_int)]


class AdamScalars(C.Structure):
    _fields_ = [("w1", C.c_float), ("beta2", C.c_float), ("w2", C.c_float), ("bc2_sqrt", C.c_float),
                ("neg_step", C.c_float), ("eps", C.c_float), ("step", C.c_int), ("pad", C.c_int)]


class GemmProblem(C.Structure):
    _fields_ = [
        ("seg", GemmSeg * 3), ("nseg", C.c_int), ("I", C.c_int), ("J", C.c_int),
        ("C", c_vp), ("c_gs", C.c_int64), ("c_ld", C.c_int), ("act", C.c_int),
        ("bias", c_vp), ("bias_gs", C.c_int64),
        ("aux", c_vp), ("aux_gs", C.c_int64), ("aux_ld", C.c_int), ("accumulate", C.c_int),
        ("P", c_vp), ("M", c_vp), ("V", c_vp), ("p_gs", C.c_int64), ("p_ld", C.c_int),
        ("noise_sigma", C.c_float), ("noise_clip", C.c_float),
        ("adam", c_vp), ("tile_begin", C.c_int), ("tiles_j", C.c_int),
    ]


# name -> (restype, argtypes); mirrors include/b200rl.h one to one
SIGNATURES = {
    "b200rl_last_error": (C.c_char_p, []),
    "b200rl_version": (C.c_int, []),
    "b200rl_is_cuda_build": (C.c_int, []),
    "b200rl_set_mma": (C.c_int, [C.c_int]),
    "b200rl_set_mega": (C.c_int, [C.c_int]),
    "b200rl_malloc": (C.c_int, [C.POINTER(c_vp), C.c_uint64]),
    "b200rl_free": (C.c_int, [c_vp]),
    "b200rl_memcpy_h2d": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp]),
    "b200rl_memcpy_d2h": (C.c_int, [c_vp, c_vp, C.c_uint64, c_vp]),
    "b200rl_stream_sync": (C.c_int, [c_vp]),
    # replay
    "b200rl_replay_create": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "b200rl_replay_destroy": (C.c_int, [c_vp]),
    "b200rl_replay_row_stride": (C.c_int, [c_vp]),
    "b200rl_replay_row_layout": (C.c_int, [c_vp, C.POINTER(C.c_int)]),
    "b200rl_replay_append": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, c_vp]),
    "b200rl_replay_append_device": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, c_vp]),
    "b200rl_replay_set_action_norm": (C.c_int, [c_vp, c_vp, c_vp]),
    "b200rl_replay_append_raw": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, c_vp]),
    "b200rl_replay_append_all": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp]),
    "b200rl_replay_append_transitions": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp]),
    "b200rl_replay_sample": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b200rl_replay_last_indices": (C.c_int, [c_vp, c_vp, c_vp]),
    "b200rl_replay_update_priority": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "b200rl_replay_reset_max_priority": (C.c_int, [c_vp, c_vp]),
    "b200rl_replay_get_state": (C.c_int, [c_vp, C.c_int, c_i64p, c_i64p, c_f32p, c_vp]),
    "b200rl_replay_set_cursor": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64]),
    "b200rl_replay_set_rng": (C.c_int, [c_vp, C.c_uint64]),
    "b200rl_replay_read_priority": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64, c_vp, c_vp]),
    "b200rl_replay_write_priority": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64, c_vp, C.c_float, c_vp]),
    "b200rl_replay_read_rows": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64, c_vp, c_vp]),
    # agents
    "b200rl_td7_create": (C.c_int, [C.POINTER(TD7Config), C.POINTER(c_vp)]),
    "b200rl_mlp_create": (C.c_int, [C.POINTER(MlpConfig), C.POINTER(c_vp)]),
    "b200rl_agent_destroy": (C.c_int, [c_vp]),
    "b200rl_agent_net_count": (C.c_int, [c_vp]),
    "b200rl_agent_net_info": (C.c_int, [c_vp, C.c_int, c_i64p, C.POINTER(C.c_int)]),
    "b200rl_agent_layer_info": (C.c_int, [c_vp, C.c_int, C.c_int, c_i64p, c_i64p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200rl_agent_read": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "b200rl_agent_write": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "b200rl_agent_copy_from": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "b200rl_agent_get_counters": (C.c_int, [c_vp, c_i64p, c_i64p]),
    "b200rl_agent_set_counters": (C.c_int, [c_vp, C.c_int64, c_i64p, c_vp]),
    "b200rl_td7_value_range": (C.c_int, [c_vp, C.c_int, c_vp, c_vp]),
    "b200rl_td7_set_value_range": (C.c_int, [c_vp, C.c_int, c_vp, c_vp]),
    "b200rl_agent_set_rng": (C.c_int, [c_vp, C.c_uint64]),
    "b200rl_agent_layer_split": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "b200rl_agent_io_ld": (C.c_int, [c_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200rl_agent_io": (C.c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), c_i64p]),
    "b200rl_replay_sample_to_agent": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "b200rl_agent_stage_batch": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b200rl_agent_update": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "b200rl_agent_update_n": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "b200rl_agent_info_cols": (C.c_int, [c_vp]),
    "b200rl_agent_wait_info": (C.c_int, [c_vp, C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "b200rl_agent_read_info": (C.c_int, [c_vp, C.c_int64, C.c_int, C.c_int, c_vp, c_vp]),
    "b200rl_agent_act": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "b200rl_agent_act_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "b200rl_agent_debug_read": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "b200rl_agent_launches_per_update": (C.c_int, [c_vp, C.c_int]),
    # ops
    "b200rl_op_gemm": (C.c_int, [C.c_int, C.POINTER(GemmProblem), C.c_int, C.c_int, c_vp]),
    "b200rl_op_gemm_timed": (C.c_int, [C.c_int, C.POINTER(GemmProblem), C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, c_vp]),
    "b200rl_replay_time_sample": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_f32p, c_vp]),
    "b200rl_op_avgl1norm": (C.c_int, [C.c_int, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    "b200rl_op_adam": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, c_vp]),
    "b200rl_op_polyak": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_double, c_vp]),
    "b200rl_op_sac_sample": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_double, C.c_double, c_vp]),
    "b200rl_op_sac_sample_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_double, C.c_double, c_vp]),
}


class B200RLError(RuntimeError):
    pass


def bind(cdll: C.CDLL) -> C.CDLL:
    """Declare argument / result types on a loaded library (raises if a symbol is missing)."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return cdll


_LIB = None


def lib() -> C.CDLL:
    """The nvcc-built library.  Fails loudly when it has not been built -- no fallback exists."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise B200RLError(
                f"{LIB_PATH} is missing: build it with `python -m sac_td3_td7_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback for this path."
            )
        cdll = bind(C.CDLL(LIB_PATH))
        if cdll.b200rl_is_cuda_build() != 1:
            raise B200RLError(f"{LIB_PATH} is not a CUDA build")
        _LIB = cdll
    return _LIB


def check(status: int, cdll: C.CDLL | None = None) -> None:
    if status != 0:
        msg = (cdll or lib()).b200rl_last_error()
        raise B200RLError(f"b200rl error {status}: {msg.decode() if msg else '?'}")


_raw_stream = None


def current_stream_ptr() -> int:
    """``cudaStream_t`` of torch's current stream on the current device, as an int.

    ``torch.cuda.current_stream().cuda_stream`` costs ~3 us per call (device-index resolution, availability checks,
    a Stream object); the per-step API makes four such calls per update, so the raw binding is used when present."""
    global _raw_stream
    if _raw_stream is None:
        import torch

        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        if raw is not None:
            get_dev = torch._C._cuda_getDevice
            _raw_stream = lambda: raw(get_dev())  # noqa: E731
        else:
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    return _raw_stream()
