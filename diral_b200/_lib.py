"""ctypes binding of libdiral_env.so (the C ABI in include/diral_env.h).

There is no CPU implementation behind this module: if the shared object is missing it is built with
nvcc, and if that is impossible the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

ABI_VERSION = 3
MODES = {"my_step": 0, "my_step_design": 1, "my_step_ch": 2}
ERR_NAMES = {-1: "DIRAL_ERR_ARG", -2: "DIRAL_ERR_CUDA", -3: "DIRAL_ERR_UNBOUND", -4: "DIRAL_ERR_SEQ_RANGE",
             -5: "DIRAL_ERR_UNSUPPORTED"}


class DiralCfg(C.Structure):
    """POD mirror of ``diral_cfg`` (include/diral_env.h)."""
    _fields_ = [("E", C.c_int64), ("env0", C.c_int64),
                ("N", C.c_int32), ("R", C.c_int32), ("B", C.c_int32),
                ("L", C.c_double), ("C", C.c_double), ("W", C.c_double),
                ("reward_design", C.c_int32), ("state_type", C.c_int32), ("toy", C.c_int32),
                ("mobility", C.c_int32), ("mobility_vary", C.c_int32), ("design_topology", C.c_int32),
                ("add_action", C.c_int32), ("action_binary", C.c_int32), ("add_channel_obs", C.c_int32),
                ("add_reward", C.c_int32), ("add_index", C.c_int32), ("add_velocity", C.c_int32),
                ("add_position", C.c_int32), ("add_positional_dist", C.c_int32), ("add_piggy", C.c_int32),
                ("pos_dist_type", C.c_int32), ("fingerprint", C.c_int32),
                ("age_threshold", C.c_int32), ("sentinel", C.c_double)]


class DiralBuffers(C.Structure):
    """POD mirror of ``diral_buffers`` (include/diral_env.h)."""
    _fields_ = [("pos_x", C.c_void_p), ("pos_y", C.c_void_p), ("vel", C.c_void_p),
                ("tab_seq", C.c_void_p), ("tab_lu", C.c_void_p), ("tab_x", C.c_void_p),
                ("lat", C.c_void_p), ("obs", C.c_void_p), ("rews", C.c_void_p), ("state", C.c_void_p),
                ("acc_reward", C.c_void_p), ("acc_count", C.c_void_p), ("scratch", C.c_void_p),
                ("trace", C.c_void_p), ("trace_len", C.c_int64), ("ring", C.c_void_p)]


class DiralShaping(C.Structure):
    """POD mirror of ``diral_shaping`` (include/diral_env.h)."""
    _fields_ = [("ia_averaging", C.c_int32), ("ia_penalty_enable", C.c_int32), ("ia_penalty_threshold", C.c_int32),
                ("global_reward_avg", C.c_int32), ("ia_penalty_value", C.c_double)]


class DiralSpsCfg(C.Structure):
    """POD mirror of ``diral_sps_cfg`` (include/diral_env.h)."""
    _fields_ = [("rssi_threshold", C.c_double), ("inc_db", C.c_double), ("prob_resource_keep", C.c_double),
                ("min_candidates", C.c_double)]


# every symbol include/diral_env.h declares: name -> (restype, argtypes)
_P, _I64, _U64, _I32, _D = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_double
SYMBOLS = {
    "diral_abi_version": (C.c_int32, []),
    "diral_last_error": (C.c_char_p, []),
    "diral_state_space": (C.c_int32, [C.POINTER(DiralCfg)]),
    "diral_state_bytes": (C.c_size_t, [C.POINTER(DiralCfg)]),
    "diral_scratch_bytes": (C.c_size_t, [C.POINTER(DiralCfg)]),
    "diral_create": (C.c_int, [C.POINTER(DiralCfg), C.POINTER(C.c_void_p)]),
    "diral_destroy": (C.c_int, [_P]),
    "diral_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "diral_bind": (C.c_int, [_P, C.POINTER(DiralBuffers)]),
    "diral_reset": (C.c_int, [_P, _P, _P, _P, _U64, _P]),
    "diral_sample": (C.c_int, [_P, _U64, _I64, _P, _P]),
    "diral_step": (C.c_int, [_P, C.c_int, _P, _I64, C.c_int, _D, _D, _U64, _P, _P]),
    "diral_obtain_state": (C.c_int, [_P, _P, _P, _P, _D, _D, _P, _P]),
    "diral_rollout": (C.c_int, [_P, C.c_int, _I32, _I64, _U64, _P]),
    "diral_update_velocity": (C.c_int, [_P, _P, _U64, _I64, _P]),
    "diral_information_age": (C.c_int, [_P, _I64, _P, _P]),
    "diral_episode_metrics": (C.c_int, [_P, _I64, _P, _P]),
    "diral_shape_rewards": (C.c_int, [_P, C.POINTER(DiralShaping), _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "diral_ring_gather": (C.c_int, [_P, _I64, _I64, _I64, _I32, _P, _I32, _I32, _P, _P]),
    "diral_wire_vpd": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _D, _I32, _P, _P]),
    "diral_sps_step": (C.c_int, [_I64, _I32, _P, C.POINTER(DiralSpsCfg), _P, _U64, _I64, _P, _P, _P, _P, _P]),
    "diral_step_host": (C.c_int, [_P, C.c_int, _P, _I64, _D, _D, _P, _P, _P, _P]),
    "diral_step_host_begin": (C.c_int, [_P, C.c_int, _P, _I64, _D, _D, _P, _P, _P]),
    "diral_step_host_wait": (C.c_int, [_P]),
    "diral_launch_count": (C.c_int64, [_P]),
    "diral_get_option": (C.c_int64, [_P, C.c_char_p]),
    "diral_reset_topology": (C.c_int, [_P, _P, _P, _P, _U64, _P]),
    "diral_expand_state_host": (C.c_int, [C.POINTER(DiralCfg), _I64, _P, _P, _P, _P, _P, _P, _P, _D, _D, _I32, _P]),
    "diral_host_trace": (C.c_int32, [_P, _P, _I32]),
    "diral_materialize_x": (C.c_int, [_P, _P, _P]),
    "diral_ring_put": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
}

_lib = None


class DiralError(RuntimeError):
    """A libdiral_env.so entry point returned a negative status."""

    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "DIRAL_ERR"), code, message))
        self.code = code


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) libdiral_env.so and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DIRAL_ENV_LIB")                   # DIRAL_ENV_LIB: tuning variants only
    if not path:
        # build() is a no-op when the library matches the sources (content hash, so a checkout or a copy to the
        # GPU box does not trigger it); it raises when a rebuild is needed and nvcc is absent: no fallback exists
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.diral_abi_version() != ABI_VERSION:
        raise ImportError("libdiral_env.so ABI %d != binding ABI %d; rebuild with python -m diral_b200._build --force"
                          % (lib.diral_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise DiralError(rc, load().diral_last_error().decode("utf-8", "replace"))
