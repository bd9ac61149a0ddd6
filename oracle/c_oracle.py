"""ctypes wrapper of oracle/ipp_oracle.c (TEST INFRASTRUCTURE / CPU baseline, not product code)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

REWARD_ENTROPY = 1
FLAG_ADAPTIVE = 4
FLAG_NO_DSIZE_QUIRK = 8
FLAG_PREDICT_ONLY = 256


class orc_cfg(C.Structure):
    _fields_ = [("x_dim", C.c_int32), ("y_dim", C.c_int32), ("cost_mode", C.c_int32), ("pad", C.c_int32)] + [
        (n, C.c_double) for n in ("res", "tan_x", "tan_y", "coeff_a", "coeff_b", "rf_alt", "max_v", "max_a", "thr", "kappa")
    ]


_lib = None


_native = False


def use_native_build() -> str:
    """Switch to a -O3 -march=native build made on THIS machine (CPU-baseline timing only; call before the first use).
    Falls back to the portable build when gcc is missing."""
    global _lib, _native
    try:
        path = _build.build_oracle(force=True, native=True)
        _lib, _native = None, True
        return f"-O3 -march=native, built on this host ({os.path.basename(path)})"
    except Exception as exc:
        return f"-O2 portable build (native build failed: {exc!r})"


def lib():
    global _lib
    if _lib is None:
        if _native:
            path = _build.LIB_NATIVE
        else:
            path = _build.LIB if os.path.exists(_build.LIB) and os.path.getmtime(_build.LIB) >= os.path.getmtime(_build.SRC) else _build.build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_step.restype = C.c_int
        _lib.orc_step.argtypes = [C.POINTER(orc_cfg), C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_uint64, C.c_int64, C.c_uint64, C.c_int,
                                                                                   C.c_void_p, C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_set_threads.argtypes = [C.c_int]
    return _lib


def make_cfg(x_dim, y_dim, resolution, angle_x=60.0, angle_y=60.0, coeff_a=0.05, coeff_b=0.2, rf_altitude=10.0, max_v=2.0, max_a=2.0,
             value_threshold=0.4, interval_factor=0.0, **_unused) -> orc_cfg:
    c = orc_cfg()
    c.x_dim, c.y_dim = int(x_dim), int(y_dim)
    c.cost_mode = 0 if max_v is None else 1
    c.res = resolution
    c.tan_x = float(np.tan(0.5 * np.radians(angle_x)))
    c.tan_y = float(np.tan(0.5 * np.radians(angle_y)))
    c.coeff_a, c.coeff_b, c.rf_alt = coeff_a, coeff_b, rf_altitude
    c.max_v, c.max_a = (max_v or 0.0), (max_a or 0.0)
    c.thr, c.kappa = value_threshold, interval_factor
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def step(cfg: orc_cfg, gt, mean, var, prev, actions, eps=None, seed=0, env_offset=0, step_idx=0, flags=0, z_out=None):
    """In-place fp64 step over B envs; returns rewards (B,)."""
    B = mean.shape[0]
    for a in (mean, var, prev, actions):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    if gt is not None:
        assert gt.dtype == np.float64 and gt.flags["C_CONTIGUOUS"]
    stride = 0
    if eps is not None:
        assert eps.dtype == np.float64 and eps.flags["C_CONTIGUOUS"]
        stride = eps.shape[1]
    if z_out is not None:
        assert z_out.dtype == np.float64 and z_out.flags["C_CONTIGUOUS"]
        stride = stride or z_out.shape[1]
        assert z_out.shape[1] == stride
    reward = np.zeros(B)
    rc = lib().orc_step(C.byref(cfg), B, _p(gt), _p(mean), _p(var), _p(prev), _p(actions), _p(eps), stride, seed, env_offset, step_idx,
                        flags, _p(reward), _p(z_out))
    if rc != 0:
        raise RuntimeError(f"orc_step returned {rc}")
    return reward


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().orc_set_threads(n)
    return n


def num_threads() -> int:
    return int(lib().orc_num_threads())
