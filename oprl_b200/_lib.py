"""ctypes binding of the C ABI declared in include/oprl_b200.h.

The CUDA library is the product: importing this module without a built
``liboprl_b200.so`` raises (there is no CPU / PyTorch fallback for the update path).
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` at the repo root.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboprl_b200.so")

ALGO = {"ddpg": 0, "td3": 1, "sac": 2, "tqc": 3}
GEMM_3XTF32, GEMM_TF32, GEMM_SIMT = 0, 1, 2
NET_ACTOR, NET_CRITIC = 0, 1
UPDATE_ACTOR = 1
GRAD_TAIL = 8  # OPRL_GRAD_TAIL
SEG_ALL, SEG_CRITIC_GRAD, SEG_CRITIC_STEP_ACTOR_GRAD, SEG_ACTOR_STEP = -1, 0, 1, 2
SCALARS = ("critic_loss", "actor_loss", "alpha_loss", "q_mean", "q_target_mean", "logpi_mean",
           "q_err_mean", "alpha")


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "algo", "state_dim", "action_dim", "actor_hidden", "actor_layers", "critic_hidden",
        "critic_layers", "n_critics", "n_quantiles", "top_quantiles_to_drop", "tune_alpha",
        "gemm_mode", "device", "world_size")] + [(n, C.c_double) for n in (
        "gamma", "tau", "lr_actor", "lr_critic", "lr_alpha", "policy_noise", "noise_clip",
        "max_action", "alpha_init", "target_entropy")] + [("seed", C.c_ulonglong)]


class State(C.Structure):
    _fields_ = [("tick", C.c_ulonglong), ("step_actor", C.c_int), ("step_critic", C.c_int),
                ("step_alpha", C.c_int), ("pad", C.c_int), ("log_alpha", C.c_double),
                ("m_alpha", C.c_double), ("v_alpha", C.c_double)]


_P = C.c_void_p
_SIGNATURES = {
    "oprl_last_error": (C.c_char_p, []),
    "oprl_abi_version": (C.c_int, []),
    "oprl_engine_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(_P)]),
    "oprl_engine_destroy": (None, [_P]),
    "oprl_engine_arena_floats": (C.c_longlong, [_P, C.c_int]),
    "oprl_engine_bind_arena": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    "oprl_engine_sync_params": (C.c_int, [_P]),
    "oprl_buffer_bind": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "oprl_buffer_set_prefix": (C.c_int, [_P, _P, C.c_int]),
    "oprl_buffer_set_nstep": (C.c_int, [_P, C.c_int, C.c_double]),
    "oprl_batch_bind": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "oprl_sample": (C.c_int, [_P, _P, C.c_int]),
    "oprl_load_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "oprl_load_batch_host": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int]),
    "oprl_set_noise": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "oprl_update": (C.c_int, [_P, C.c_int, C.c_int]),
    "oprl_step": (C.c_int, [_P, C.c_int, C.c_int]),
    "oprl_get_scalars": (C.c_int, [_P, _P, C.c_int]),
    "oprl_scalars_enqueue": (C.c_int, [_P]),
    "oprl_scalars_wait": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "oprl_get_state": (C.c_int, [_P, C.POINTER(State)]),
    "oprl_set_state": (C.c_int, [_P, C.POINTER(State)]),
    "oprl_sync": (C.c_int, [_P]),
    "oprl_stream": (_P, [_P]),
    "oprl_engine_set_stream": (C.c_int, [_P, _P]),
    "oprl_engine_set_world_size": (C.c_int, [_P, C.c_int]),
    "oprl_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "oprl_comm_connect": (C.c_int, [_P, _P, _P]),
    "oprl_update_launches": (C.c_int, [_P, C.c_int, C.c_int]),
    "oprl_profile": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "oprl_scatter_transitions": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "oprl_chain_prof": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "oprl_gather_rows": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                   C.c_int, _P, _P, _P, _P, _P, _P]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load liboprl_b200.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} is missing: the CUDA engine is not built and oprl_b200 has no "
                "CPU fallback.  Run `python -c 'import __graft_entry__ as g; g.build()'`.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise EngineError(lib().oprl_last_error().decode() or f"oprl_b200 error {rc}")
    return rc
