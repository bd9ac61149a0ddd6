"""ctypes binding of the C ABI in ``include/ipp_b200.h`` (``csrc/libipp_b200.so``).

This is the only place the shared library is loaded.  There is no CPU fallback: if the library
is missing or a CUDA call fails, an exception is raised (``IppLibraryError`` / ``IppError``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPP_B200_LIB") or os.path.join(_HERE, "csrc", "libipp_b200.so")  # override: ablation builds

IPP_ABI_VERSION = 1
IPP_OK = 0
IPP_ERR_INVALID = -1
IPP_ERR_CUDA = -2
IPP_ERR_NOMEM = -3
IPP_ERR_UNSUPPORTED = -4

REWARD_TRACE = 0
REWARD_GAUSS_ENTROPY = 1
FLAG_ADAPTIVE = 4
FLAG_NO_DSIZE_QUIRK = 8
FLAG_LOGODDS = 16
FLAG_NO_COMMIT = 32
FLAG_KEEP_PREV = 64
OBS_COSTS = 256

LAYOUT_PLANES = 0
LAYOUT_MV = 1
LAYOUT_TILED = 2
LAYOUT_SUPER = 3
LAYOUT_SPLIT = 4
LAYOUT_NAMES = {"planes": LAYOUT_PLANES, "mv": LAYOUT_MV, "tiled": LAYOUT_TILED, "super": LAYOUT_SUPER, "split": LAYOUT_SPLIT}
FIELD_HOTSPOT, FIELD_SPLIT = 1, 2
COST_DISTANCE = 0
COST_FLIGHT_TIME = 1
MAX_ALTITUDE_LEVELS = 32
NUM_METRICS = 8

PTR_MEAN, PTR_VAR, PTR_GT, PTR_REWARD, PTR_STREAM = range(5)
PATH_LSU, PATH_ASYNC = 0, 1
OPT_STEP_PATH = 1
OPT_LAUNCHES_LSU, OPT_LAUNCHES_ASYNC = 2, 3
OPT_ZERO_COPY, OPT_ZERO_COPY_STEPS, OPT_IDS_FETCH_STEPS = 4, 5, 6
ZERO_COPY_REWARDS, ZERO_COPY_IDS, ZERO_COPY_IDS_FETCH = 1, 2, 4
STEP_SLOTS = 2


class IppLibraryError(RuntimeError):
    """The CUDA extension is missing / not loadable.  The product path never falls back to CPU."""


class IppError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ipp_b200 error {code}: {message}")
        self.code = code


class ipp_config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("abi_version", C.c_uint32),
        ("device", C.c_int32),
        ("batch", C.c_int32),
        ("x_dim", C.c_int32),
        ("y_dim", C.c_int32),
        ("layout", C.c_int32),
        ("cost_mode", C.c_int32),
        ("resolution", C.c_double),
        ("angle_x_deg", C.c_double),
        ("angle_y_deg", C.c_double),
        ("tan_half_x", C.c_double),
        ("tan_half_y", C.c_double),
        ("coeff_a", C.c_double),
        ("coeff_b", C.c_double),
        ("rf_altitude", C.c_double),
        ("min_altitude", C.c_double),
        ("max_altitude", C.c_double),
        ("altitude_spacing", C.c_double),
        ("max_v", C.c_double),
        ("max_a", C.c_double),
        ("value_threshold", C.c_double),
        ("interval_factor", C.c_double),
        ("seed", C.c_uint64),
        ("env_id_offset", C.c_int64),
        ("stream", C.c_void_p),
    ]


class ipp_info(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("x_dim", C.c_int32),
        ("y_dim", C.c_int32),
        ("layout", C.c_int32),
        ("num_altitude_levels", C.c_int32),
        ("num_actions", C.c_int32),
        ("max_measurements", C.c_int32),
        ("sm_count", C.c_int32),
        ("launches", C.c_uint64),
        ("steps", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("altitude", C.c_double * MAX_ALTITUDE_LEVELS),
        ("radius_x", C.c_int32 * MAX_ALTITUDE_LEVELS),
        ("radius_y", C.c_int32 * MAX_ALTITUDE_LEVELS),
    ]


class ipp_mcts_config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("n_trees", C.c_int32),
        ("first_env", C.c_int32),
        ("num_simulations", C.c_int32),
        ("episode_horizon", C.c_int32),
        ("step_flags", C.c_uint32),
        ("puct_init", C.c_double),
        ("puct_base", C.c_double),
        ("gamma", C.c_double),
        ("forced_playout_factor", C.c_double),
        ("max_valid_action_distance", C.c_double),
        ("dirichlet_eps", C.c_double),
    ]


class ipp_mcts_info(C.Structure):
    _fields_ = [
        ("n_trees", C.c_int32),
        ("max_nodes", C.c_int32),
        ("levels", C.c_int32),
        ("window_dim", C.c_int32),
        ("window_radius", C.c_int32),
        ("window_slots", C.c_int32),
        ("max_path", C.c_int32),
        ("simulations", C.c_int32),
        ("device_bytes", C.c_uint64),
        ("launches", C.c_uint64),
        ("edges", C.c_uint64),
    ]


class ipp_ring_config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("device", C.c_int32),
        ("capacity", C.c_int64),
        ("channels", C.c_int32),
        ("y_dim", C.c_int32),
        ("x_dim", C.c_int32),
        ("policy_slots", C.c_int32),
        ("stream", C.c_void_p),
    ]


class ipp_ring_info(C.Structure):
    _fields_ = [
        ("capacity", C.c_int64),
        ("size", C.c_int64),
        ("head", C.c_int64),
        ("pushed", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("launches", C.c_uint64),
    ]


(RING_PTR_OBS, RING_PTR_POLICY, RING_PTR_MASK, RING_PTR_VALUE, RING_PTR_REWARD, RING_PTR_PRIORITY, RING_PTR_LAST_INDICES,
 RING_PTR_LAST_WEIGHTS, RING_PTR_STREAM) = range(9)

MCTS_MAX_PATH = 8
MCTS_LEAF_TERMINAL, MCTS_LEAF_EVAL = 0, 1
MCTS_LEAF_WORDS = 8
MCTS_PTR_LEAF_INFO, MCTS_PTR_PATH_ACTIONS, MCTS_PTR_PATH_REWARDS = range(3)

_P = C.c_void_p
_I32, _U32, _F32 = C.c_int32, C.c_uint32, C.c_float

# name -> (restype, argtypes); must list every symbol include/ipp_b200.h declares
SIGNATURES = {
    "ipp_create": (C.c_int, [C.POINTER(ipp_config), C.POINTER(_P)]),
    "ipp_destroy": (None, [_P]),
    "ipp_last_error": (C.c_char_p, [_P]),
    "ipp_get_info": (C.c_int, [_P, C.POINTER(ipp_info)]),
    "ipp_sync": (C.c_int, [_P]),
    "ipp_reset": (C.c_int, [_P, _F32, _F32, _P, _P]),
    "ipp_set_ground_truth": (C.c_int, [_P, _P, _I32, _I32, _I32]),
    "ipp_get_ground_truth": (C.c_int, [_P, _P, _I32, _I32, _I32]),
    "ipp_synth_ground_truth": (C.c_int, [_P, C.c_uint64]),
    "ipp_generate_field": (C.c_int, [_P, _I32, _I32, C.c_uint64, _I32, _I32]),
    "ipp_reset_shuffled": (C.c_int, [_P, _F32, _I32, _F32, C.c_uint64, _P]),
    "ipp_generate_ground_truth": (C.c_int, [_P, C.c_double, C.c_uint64, _P, _I32, _I32]),
    "ipp_get_state": (C.c_int, [_P, _P, _P, _I32, _I32, _I32]),
    "ipp_set_state": (C.c_int, [_P, _P, _P, _I32, _I32, _I32]),
    "ipp_set_prev_pose": (C.c_int, [_P, _P]),
    "ipp_get_prev_pose": (C.c_int, [_P, _P]),
    "ipp_step": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _U32]),
    "ipp_step_device": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _U32]),
    "ipp_step_submit": (C.c_int, [_P, _I32, _P, _P, _U32]),
    "ipp_step_wait": (C.c_int, [_P, _I32]),
    "ipp_measure": (C.c_int, [_P, _P, _P, _P, _I32, _P, _U32]),
    "ipp_update": (C.c_int, [_P, _P, _P, _P, _I32, _P, _U32]),
    "ipp_predict": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _U32]),
    "ipp_predict_device": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _U32]),
    "ipp_rollout": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _U32]),
    "ipp_rollout_device": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _U32]),
    "ipp_observe": (C.c_int, [_P, _I32, _I32, _P, _P, _U32, _P, _I32]),
    "ipp_eval": (C.c_int, [_P, _P]),
    "ipp_eval_device": (C.c_int, [_P, _P]),
    "ipp_device_ptr": (_P, [_P, _I32]),
    "ipp_set_option": (C.c_int, [_P, _I32, C.c_int64]),
    "ipp_get_option": (C.c_int64, [_P, _I32]),
    "ipp_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "ipp_host_free": (C.c_int, [_P]),
    "ipp_kalman_blocks": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    # include/ipp_mcts.h
    "ipp_mcts_create": (C.c_int, [_P, C.POINTER(ipp_mcts_config), C.POINTER(_P)]),
    "ipp_mcts_destroy": (None, [_P]),
    "ipp_mcts_last_error": (C.c_char_p, [_P]),
    "ipp_mcts_get_info": (C.c_int, [_P, C.POINTER(ipp_mcts_info)]),
    "ipp_mcts_begin": (C.c_int, [_P, _P, _P]),
    "ipp_mcts_simulate_begin": (C.c_int, [_P, _P]),
    "ipp_mcts_simulate_end": (C.c_int, [_P, _P, _P, _P, _P, _I32]),
    "ipp_mcts_root_stats": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ipp_mcts_get_paths": (C.c_int, [_P, _P, _P]),
    "ipp_mcts_device_ptr": (_P, [_P, _I32]),
    # include/ipp_experience.h
    "ipp_ring_create": (C.c_int, [C.POINTER(ipp_ring_config), C.POINTER(_P)]),
    "ipp_ring_destroy": (None, [_P]),
    "ipp_ring_last_error": (C.c_char_p, [_P]),
    "ipp_ring_get_info": (C.c_int, [_P, C.POINTER(ipp_ring_info)]),
    "ipp_ring_value_targets": (C.c_int, [_P, _P, _P, _I32, _I32, C.c_double, _I32, _P, _P, _I32]),
    "ipp_ring_push": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _F32, _I32]),
    "ipp_ring_reset_priorities": (C.c_int, [_P]),
    "ipp_ring_sample": (C.c_int, [_P, _I32, C.c_double, C.c_double, _P, C.c_uint64, _P, _P, _I32]),
    "ipp_ring_gather": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _P, _I32]),
    "ipp_ring_update_priorities": (C.c_int, [_P, _I32, _P, _P, _I32]),
    "ipp_ring_get_priorities": (C.c_int, [_P, _P]),
    "ipp_ring_device_ptr": (_P, [_P, _I32]),
}

_lib: Optional[C.CDLL] = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load ``libipp_b200.so`` and bind every exported symbol.  Raises IppLibraryError."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise IppLibraryError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback."
        )
    try:
        lib = C.CDLL(p)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise IppLibraryError(f"cannot load {p}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise IppLibraryError(f"{p} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, handle, code: int) -> None:
    if code != IPP_OK:
        msg = lib.ipp_last_error(handle)
        raise IppError(code, msg.decode() if msg else "unknown error")
