"""Batched IPP environment engine — Python host of the sm_100a CUDA path.

``BatchedEngine`` owns ``batch`` independent env instances resident in HBM and exposes the hot
path of the reference as batched calls:

=====================  ============================================================================
``step``               take_measurement + update_grid_map + reward, one fused kernel
                       (simulations/simulations.py:26-34, mapping/mappings.py:114-215,
                       planning/common/rewards.py:8-31)
``predict``            ``simulate_prediction_step`` (planning/common/optimization.py:14-30)
``measure`` / ``update``  the two halves of ``step`` (sensors/cameras.py:108-116,
                       mapping/mappings.py:114-153)
``eval``               ``Mission.eval`` metrics (planning/missions.py:176-203)
``reset``              ``Mapping.init_priors`` diagonal restriction (mapping/mappings.py:217-261)
=====================  ============================================================================

All compute happens in ``csrc/libipp_b200.so`` through the C ABI (``include/ipp_b200.h``); this
module only marshals NumPy arrays.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import _capi as capi


@dataclass
class EngineConfig:
    """Hot-path configuration = the reference YAML keys the step consumes (config/example.yaml)."""

    batch: int = 1
    x_dim: int = 10
    y_dim: int = 10
    resolution: float = 4.0
    angle_x: float = 60.0
    angle_y: float = 60.0
    coeff_a: float = 0.05
    coeff_b: float = 0.2
    rf_altitude: float = 10.0
    min_altitude: float = 8.0
    max_altitude: float = 14.0
    altitude_spacing: float = 6.0
    max_v: Optional[float] = 2.0
    max_a: Optional[float] = 2.0
    value_threshold: float = 0.4
    interval_factor: float = 0.0
    layout: int = capi.LAYOUT_PLANES
    device: int = 0
    seed: int = 20260925
    env_id_offset: int = 0
    stream: Optional[int] = None

    @classmethod
    def from_params(cls, params: Dict, batch: int = 1, **overrides) -> "EngineConfig":
        """Read the same nested dict the reference factories read.  Missing required keys raise
        ValueError like the reference (e.g. mapping/grid_maps.py:13-24, sensor_factories.py:27-47)."""
        try:
            env = params["environment"]
            sen = params["sensor"]
            fov = sen["field_of_view"]
            model = sen["model"]
            kw = dict(
                batch=batch,
                x_dim=int(env["x_dim"]),
                y_dim=int(env["y_dim"]),
                resolution=float(env["resolution"]),
                angle_x=float(fov["angle_x"]),
                angle_y=float(fov["angle_y"]),
                coeff_a=float(model["coeff_a"]),
                coeff_b=float(model["coeff_b"]),
            )
        except KeyError as exc:
            raise ValueError(f"Cannot find {exc} specification in config file!") from exc
        exp = params.get("experiment", {})
        con = exp.get("constraints", {})
        sce = exp.get("scenario", {})
        for k in ("min_altitude", "max_altitude", "altitude_spacing"):
            if k in con:
                kw[k] = float(con[k])
        for k in ("value_threshold", "interval_factor"):
            if k in sce:
                kw[k] = float(sce[k])
        uav = exp.get("uav")
        if uav is not None:
            kw["max_v"], kw["max_a"] = float(uav["max_v"]), float(uav["max_a"])
        else:  # uav_specifications=None in the reference: costs are Euclidean distances (planning/common/actions.py:8-16)
            kw["max_v"] = kw["max_a"] = None
        backend = params.get("mapping", {}).get("b200", {})
        if "layout" in backend:
            kw["layout"] = capi.LAYOUT_NAMES[backend["layout"]]
        kw.update(overrides)
        return cls(**kw)

    def to_c(self) -> capi.ipp_config:
        c = capi.ipp_config()
        c.struct_bytes = C.sizeof(capi.ipp_config)
        c.abi_version = capi.IPP_ABI_VERSION
        c.device, c.batch, c.x_dim, c.y_dim = self.device, self.batch, self.x_dim, self.y_dim
        c.layout = self.layout
        c.cost_mode = capi.COST_DISTANCE if self.max_v is None else capi.COST_FLIGHT_TIME
        c.resolution = self.resolution
        c.angle_x_deg, c.angle_y_deg = self.angle_x, self.angle_y
        # NumPy's own value, same expression as sensors/cameras.py:44-45, so floor() never flips
        c.tan_half_x = float(np.tan(0.5 * np.radians(self.angle_x)))
        c.tan_half_y = float(np.tan(0.5 * np.radians(self.angle_y)))
        c.coeff_a, c.coeff_b, c.rf_altitude = self.coeff_a, self.coeff_b, self.rf_altitude
        c.min_altitude, c.max_altitude, c.altitude_spacing = self.min_altitude, self.max_altitude, self.altitude_spacing
        c.max_v = 0.0 if self.max_v is None else self.max_v
        c.max_a = 0.0 if self.max_a is None else self.max_a
        c.value_threshold, c.interval_factor = self.value_threshold, self.interval_factor
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.env_id_offset = self.env_id_offset
        c.stream = C.c_void_p(self.stream) if self.stream else None
        return c


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PtrCache:
    """ctypes pointers of arrays that are passed again and again (``ndarray.ctypes`` costs ~2 us per access, a tenth of
    the host side of a step).  Entries are validated by identity through a weak reference."""

    def __init__(self, capacity: int = 256):
        self._d, self._cap = {}, capacity

    def __call__(self, a: Optional[np.ndarray]):
        if a is None:
            return None
        hit = self._d.get(id(a))
        if hit is not None and hit[0]() is a:
            return hit[1]
        p = a.ctypes.data_as(C.c_void_p)
        if len(self._d) >= self._cap:
            self._d.clear()
        try:
            self._d[id(a)] = (weakref.ref(a), p)
        except TypeError:
            pass
        return p


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def pinned_array(shape, dtype) -> np.ndarray:
    """A NumPy array on page-locked, mapped host memory (``ipp_host_alloc``): ``step`` reads ids from / writes rewards to such
    buffers without stream copies (IPP_OPT_ZERO_COPY).  The memory lives until the process exits or ``free_pinned(a)``."""
    lib = capi.load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    rc = lib.ipp_host_alloc(C.byref(ptr), max(n, 1))
    if rc != capi.IPP_OK:
        raise capi.IppError(rc, "ipp_host_alloc failed")
    buf = (C.c_char * max(n, 1)).from_address(ptr.value)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[a.__array_interface__["data"][0]] = ptr
    return a


def free_pinned(a: np.ndarray) -> None:
    ptr = _PINNED.pop(a.__array_interface__["data"][0], None)
    if ptr is not None:
        capi.load_library().ipp_host_free(ptr)


_PINNED = {}


class BatchedEngine:
    """``batch`` env instances on one B200; see module docstring."""

    def __init__(self, cfg: EngineConfig):
        self.cfg = cfg
        self._lib = capi.load_library()
        self._h = C.c_void_p()
        self._cptr = _PtrCache()
        ccfg = cfg.to_c()
        rc = self._lib.ipp_create(C.byref(ccfg), C.byref(self._h))
        if rc != capi.IPP_OK:
            msg = self._lib.ipp_last_error(None)
            self._h = C.c_void_p()
            raise capi.IppError(rc, msg.decode() if msg else "ipp_create failed")
        self.info = self._get_info()
        self.batch = cfg.batch
        self.x_dim, self.y_dim = cfg.x_dim, cfg.y_dim
        self.max_measurements = int(self.info.max_measurements)
        self.num_actions = int(self.info.num_actions)
        self.altitudes = np.array(self.info.altitude[: self.info.num_altitude_levels])

    # -- life cycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ipp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int) -> None:
        capi.check(self._lib, self._h, rc)

    def _get_info(self) -> capi.ipp_info:
        info = capi.ipp_info()
        self._ck(self._lib.ipp_get_info(self._h, C.byref(info)))
        return info

    @property
    def launches(self) -> int:
        return int(self._get_info().launches)

    @property
    def device_bytes(self) -> int:
        return int(self._get_info().device_bytes)

    def sync(self) -> None:
        self._ck(self._lib.ipp_sync(self._h))

    # -- kernel path selection ------------------------------------------------------------------
    PATHS = {"lsu": capi.PATH_LSU, "async": capi.PATH_ASYNC}

    def set_step_path(self, path: str) -> None:
        """Request the step kernel: "async" (default, cp.async-staged persistent kernel), or "lsu" (general
        gather kernel).  Unavailable paths fall back (see ``step_path``)."""
        self._ck(self._lib.ipp_set_option(self._h, capi.OPT_STEP_PATH, self.PATHS[path]))

    @property
    def step_path(self) -> str:
        v = self._lib.ipp_get_option(self._h, capi.OPT_STEP_PATH)
        return {b: a for a, b in self.PATHS.items()}[int(v)]

    def path_launches(self, path: str) -> int:
        return int(self._lib.ipp_get_option(self._h, capi.OPT_LAUNCHES_LSU + self.PATHS[path]))

    def set_zero_copy(self, rewards: bool = True, ids: bool = False, ids_fetch: bool = True) -> None:
        """Which pinned+mapped host buffers of ``step`` the kernel accesses in place (no stream copies around the
        launch): rewards written straight to the caller's buffer; action ids read from it (``ids``), or pulled by the
        persistent kernel itself into device memory under its first footprints (``ids_fetch``).  Pageable buffers
        always take the copy path."""
        mask = (capi.ZERO_COPY_REWARDS if rewards else 0) | (capi.ZERO_COPY_IDS if ids else 0) | (capi.ZERO_COPY_IDS_FETCH if ids_fetch else 0)
        self._ck(self._lib.ipp_set_option(self._h, capi.OPT_ZERO_COPY, mask))

    @property
    def zero_copy_steps(self) -> int:
        return int(self._lib.ipp_get_option(self._h, capi.OPT_ZERO_COPY_STEPS))

    @property
    def ids_fetch_steps(self) -> int:
        """steps whose action ids the kernel fetched from the caller's pinned buffer itself"""
        return int(self._lib.ipp_get_option(self._h, capi.OPT_IDS_FETCH_STEPS))

    # -- reset / world -------------------------------------------------------------------------
    def reset(self, prior_mean: float = 0.5, prior_var: float = 1.82, prior_var_per_env=None, init_pose=None) -> None:
        pv = None if prior_var_per_env is None else _f32(prior_var_per_env, (self.batch,))
        ip = None if init_pose is None else np.ascontiguousarray(init_pose, dtype=np.float64).reshape(3)
        self._ck(self._lib.ipp_reset(self._h, prior_mean, prior_var, _ptr(pv), _ptr(ip)))

    def set_ground_truth(self, gt, first_env: int = 0) -> None:
        gt = _f32(gt)
        if gt.ndim == 2:
            gt = gt[None]
        if gt.shape[1:] != (self.y_dim, self.x_dim):
            raise ValueError(f"ground truth must be (n, {self.y_dim}, {self.x_dim}), got {gt.shape}")
        self._ck(self._lib.ipp_set_ground_truth(self._h, _ptr(gt), first_env, gt.shape[0], 0))

    def synth_ground_truth(self, seed: int = 0) -> None:
        self._ck(self._lib.ipp_synth_ground_truth(self._h, seed))

    def generate_ground_truth(self, cluster_radius: float, seed: int = 0, white_noise=None, first_env: int = 0,
                              n_env: Optional[int] = None) -> None:
        """Gaussian random fields generated on the device (GaussianRandomField.create_ground_truth_map,
        simulations/simulations.py:43-48).  ``white_noise`` (n, y_dim, x_dim) standard normals = the reference's
        ``np.random.normal`` draw (parity mode); None = device Philox stream keyed by (seed, global env id)."""
        n = self.batch - first_env if n_env is None else n_env
        wn = None
        if white_noise is not None:
            wn = _f32(white_noise).reshape(-1, self.y_dim, self.x_dim)
            if wn.shape[0] != n:
                raise ValueError(f"white_noise must hold {n} maps")
        self._ck(self._lib.ipp_generate_ground_truth(self._h, float(cluster_radius), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(wn), first_env, n))

    FIELDS = {"hotspot_random_field": capi.FIELD_HOTSPOT, "split_random_field": capi.FIELD_SPLIT}

    def generate_field(self, kind: str, cluster_radius: int, seed: int = 0, first_env: int = 0, n_env: Optional[int] = None) -> None:
        """Piecewise-constant ground truths generated on the device (HotspotRandomField / SplitRandomField,
        simulations/simulations.py:51-123); ``kind`` is the reference's simulation type string."""
        n = self.batch - first_env if n_env is None else n_env
        self._ck(self._lib.ipp_generate_field(self._h, self.FIELDS[kind], int(cluster_radius), int(seed) & 0xFFFFFFFFFFFFFFFF, first_env, n))

    def reset_shuffled(self, prior_mean: float = 0.5, fit_gaussian_process: bool = True, scale: float = 1.82, seed: int = 0, init_pose=None) -> None:
        """``Mapping.init_priors(shuffle_prior_cov=True)`` for every env, drawn on the device (mapping/mappings.py:219-240):
        ``scale`` = signal_variance (GP mode) or prior_cov_mean."""
        ip = None if init_pose is None else np.ascontiguousarray(init_pose, dtype=np.float64).reshape(3)
        self._ck(self._lib.ipp_reset_shuffled(self._h, prior_mean, 1 if fit_gaussian_process else 0, float(scale), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                              _ptr(ip)))

    def get_ground_truth(self, first_env: int = 0, n_env: Optional[int] = None) -> np.ndarray:
        n = self.batch - first_env if n_env is None else n_env
        out = np.empty((n, self.y_dim, self.x_dim), np.float32)
        self._ck(self._lib.ipp_get_ground_truth(self._h, _ptr(out), first_env, n, 0))
        return out

    def get_state(self, first_env: int = 0, n_env: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        n = self.batch - first_env if n_env is None else n_env
        mean = np.empty((n, self.y_dim, self.x_dim), np.float32)
        var = np.empty_like(mean)
        self._ck(self._lib.ipp_get_state(self._h, _ptr(mean), _ptr(var), first_env, n, 0))
        return mean, var

    def set_state(self, mean=None, var=None, first_env: int = 0) -> None:
        n = None
        if mean is not None:
            mean = _f32(mean).reshape(-1, self.y_dim, self.x_dim)
            n = mean.shape[0]
        if var is not None:
            var = _f32(var).reshape(-1, self.y_dim, self.x_dim)
            if n is not None and var.shape[0] != n:
                raise ValueError("mean / var env counts differ")
            n = var.shape[0]
        if n is None:
            return
        self._ck(self._lib.ipp_set_state(self._h, _ptr(mean), _ptr(var), first_env, n, 0))

    def set_prev_pose(self, poses) -> None:
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(poses, np.float64), (self.batch, 3)))
        self._ck(self._lib.ipp_set_prev_pose(self._h, _ptr(p)))

    def get_prev_pose(self) -> np.ndarray:
        p = np.empty((self.batch, 3), np.float64)
        self._ck(self._lib.ipp_get_prev_pose(self._h, _ptr(p)))
        return p

    # -- hot path ---------------------------------------------------------------------------------
    @staticmethod
    def _split_actions(actions, n: int):
        a = np.asarray(actions)
        if a.ndim == 1 and np.issubdtype(a.dtype, np.integer):
            ids = np.ascontiguousarray(a, dtype=np.int32)
            if ids.shape != (n,):
                raise ValueError(f"action ids must have shape ({n},)")
            return ids, None
        poses = np.ascontiguousarray(a, dtype=np.float64)
        if poses.shape != (n, 3):
            raise ValueError(f"poses must have shape ({n}, 3)")
        return None, poses

    @staticmethod
    def _flags(reward_mode=capi.REWARD_TRACE, adaptive=False, dsize_quirk=True, logodds=False, commit=True, keep_prev=False) -> int:
        f = int(reward_mode) & 3
        if adaptive:
            f |= capi.FLAG_ADAPTIVE
        if not dsize_quirk:
            f |= capi.FLAG_NO_DSIZE_QUIRK
        if logodds:
            f |= capi.FLAG_LOGODDS
        if not commit:
            f |= capi.FLAG_NO_COMMIT
        if keep_prev:
            f |= capi.FLAG_KEEP_PREV
        return f

    def _noise(self, noise):
        if noise is None:
            return None, self.max_measurements
        nz = _f32(noise)
        if nz.ndim != 2 or nz.shape[0] != self.batch or nz.shape[1] < self.max_measurements:
            raise ValueError(f"noise must be (batch, >= {self.max_measurements}) standard normals")
        return nz, nz.shape[1]

    def step(self, actions, noise=None, reward_mode=capi.REWARD_TRACE, adaptive=False, dsize_quirk=True, logodds=False,
             return_measurements=False, out: Optional[np.ndarray] = None):
        """One executed step for every env.  ``actions``: int ids (batch,) or fp64 poses (batch, 3).
        ``noise``: (batch, >= max_measurements) standard normals, or None for the device Philox
        stream.  Returns rewards (batch,) float32 [and measurements (batch, stride)]."""
        if (type(actions) is np.ndarray and actions.dtype == np.int32 and actions.ndim == 1 and actions.shape[0] == self.batch
                and actions.flags.c_contiguous):
            ids, poses = actions, None  # the common call: no conversion, no copy
        else:
            ids, poses = self._split_actions(actions, self.batch)
        nz, stride = self._noise(noise)
        reward = out if out is not None else np.empty(self.batch, np.float32)
        z = np.zeros((self.batch, stride), np.float32) if return_measurements else None
        fl = self._flags(reward_mode, adaptive, dsize_quirk, logodds)
        cp = self._cptr
        self._ck(self._lib.ipp_step(self._h, cp(ids), _ptr(poses), _ptr(nz), stride, cp(reward), _ptr(z), fl))
        return (reward, z) if return_measurements else reward

    def step_submit(self, slot: int, action_ids: np.ndarray, out: np.ndarray, reward_mode=capi.REWARD_TRACE, adaptive=False) -> None:
        """Queue one step (``ipp_step_submit``) and return at once: ``action_ids`` int32 (batch,) and ``out`` float32 (batch,)
        must stay untouched until ``step_wait(slot)``; use pinned arrays (``torch.Tensor.pin_memory().numpy()``) so that the
        upload overlaps the running kernel and the rewards land in ``out`` without a copy.  Two slots: submit step t+1
        before waiting for step t."""
        if action_ids.dtype != np.int32 or action_ids.shape != (self.batch,) or not action_ids.flags.c_contiguous:
            raise ValueError(f"action_ids must be a contiguous int32 array of shape ({self.batch},)")
        if out.dtype != np.float32 or out.shape != (self.batch,) or not out.flags.c_contiguous:
            raise ValueError(f"out must be a contiguous float32 array of shape ({self.batch},)")
        cp = self._cptr
        self._ck(self._lib.ipp_step_submit(self._h, slot, cp(action_ids), cp(out), self._flags(reward_mode, adaptive)))

    def step_wait(self, slot: int) -> None:
        self._ck(self._lib.ipp_step_wait(self._h, slot))

    def measure(self, actions, noise=None, dsize_quirk=True) -> np.ndarray:
        ids, poses = self._split_actions(actions, self.batch)
        nz, stride = self._noise(noise)
        z = np.zeros((self.batch, stride), np.float32)
        self._ck(self._lib.ipp_measure(self._h, _ptr(ids), _ptr(poses), _ptr(nz), stride, _ptr(z), self._flags(dsize_quirk=dsize_quirk)))
        return z

    def update(self, actions, measurements, reward_mode=capi.REWARD_TRACE, adaptive=False, logodds=False, keep_prev=False) -> np.ndarray:
        ids, poses = self._split_actions(actions, self.batch)
        z = _f32(measurements)
        if z.ndim != 2 or z.shape[0] != self.batch or z.shape[1] < self.max_measurements:
            raise ValueError(f"measurements must be (batch, >= {self.max_measurements})")
        reward = np.empty(self.batch, np.float32)
        fl = self._flags(reward_mode, adaptive, logodds=logodds, keep_prev=keep_prev)
        self._ck(self._lib.ipp_update(self._h, _ptr(ids), _ptr(poses), _ptr(z), z.shape[1], _ptr(reward), fl))
        return reward

    def predict(self, actions, env_index=None, prev_poses=None, commit=True, reward_mode=capi.REWARD_TRACE, adaptive=False) -> np.ndarray:
        """Batched ``simulate_prediction_step``: rewards of the given jobs; ``commit=False`` leaves
        the variance untouched (predict_only contract), ``commit=True`` advances it."""
        a = np.asarray(actions)
        n = a.shape[0]
        ids, poses = self._split_actions(a, n)
        ei = None if env_index is None else np.ascontiguousarray(env_index, dtype=np.int32)
        if ei is not None and ei.shape != (n,):
            raise ValueError("env_index must have one entry per job")
        pp = None if prev_poses is None else np.ascontiguousarray(np.broadcast_to(np.asarray(prev_poses, np.float64), (n, 3)))
        reward = np.empty(n, np.float32)
        fl = self._flags(reward_mode, adaptive, commit=commit)
        self._ck(self._lib.ipp_predict(self._h, n, _ptr(ei), _ptr(ids), _ptr(poses), _ptr(pp), _ptr(reward), fl))
        return reward

    def rollout(self, path_actions, env_index=None, prev_poses=None, reward_mode=capi.REWARD_TRACE, adaptive=False) -> np.ndarray:
        """Path rollouts: ``path_actions`` (n_jobs, horizon) action ids (negative = end of path) replayed as chained
        ``simulate_prediction_step`` calls from each env's current belief, WITHOUT changing it (what MCTS.simulate does
        along one tree path, planning/mcts_zero/mcts.py:239-246).  Returns per-step rewards (n_jobs, horizon)."""
        pa = np.ascontiguousarray(path_actions, dtype=np.int32)
        if pa.ndim != 2:
            raise ValueError("path_actions must be (n_jobs, horizon)")
        n, h = pa.shape
        ei = None if env_index is None else np.ascontiguousarray(env_index, dtype=np.int32)
        if ei is not None and ei.shape != (n,):
            raise ValueError("env_index must have one entry per job")
        pp = None if prev_poses is None else np.ascontiguousarray(np.broadcast_to(np.asarray(prev_poses, np.float64), (n, 3)))
        rewards = np.empty((n, h), np.float32)
        fl = self._flags(reward_mode, adaptive)
        self._ck(self._lib.ipp_rollout(self._h, n, h, _ptr(ei), _ptr(pa), _ptr(pp), _ptr(rewards), fl))
        return rewards

    def observe(self, first_env: int = 0, n_env: Optional[int] = None, poses=None, budget_ratio=None, adaptive=False,
                action_costs=True) -> np.ndarray:
        """Network-input planes of the current belief, (n, C, y_dim, x_dim) fp32: the per-cell restriction of
        ``generate_input_feature_planes`` (planning/common/features.py:83-151) for one history entry —
        [variance / max, x, y, z position, budget ratio] (+ the action-cost plane)."""
        n = self.batch - first_env if n_env is None else n_env
        c = 6 if action_costs else 5
        pp = None if poses is None else np.ascontiguousarray(np.broadcast_to(np.asarray(poses, np.float64), (n, 3)))
        br = None if budget_ratio is None else np.ascontiguousarray(np.broadcast_to(np.asarray(budget_ratio, np.float32), (n,)))
        out = np.empty((n, c, self.y_dim, self.x_dim), np.float32)
        fl = (capi.FLAG_ADAPTIVE if adaptive else 0) | (capi.OBS_COSTS if action_costs else 0)
        self._ck(self._lib.ipp_observe(self._h, first_env, n, _ptr(pp), _ptr(br), fl, _ptr(out), 0))
        return out

    def eval(self) -> np.ndarray:
        out = np.empty((self.batch, capi.NUM_METRICS), np.float32)
        self._ck(self._lib.ipp_eval(self._h, _ptr(out)))
        return out

    # -- device-pointer (zero-copy) path ------------------------------------------------------------
    def device_ptr(self, which: int) -> int:
        p = self._lib.ipp_device_ptr(self._h, which)
        return int(p) if p else 0

    @property
    def stream(self) -> int:
        return self.device_ptr(capi.PTR_STREAM)

    def step_device(self, action_ids_ptr: int = 0, poses_ptr: int = 0, noise_ptr: int = 0, noise_stride: int = 0, reward_ptr: int = 0,
                    reward_mode=capi.REWARD_TRACE, adaptive=False, logodds=False) -> None:
        """Asynchronous step on device-resident inputs/outputs (raw device addresses)."""
        fl = self._flags(reward_mode, adaptive, logodds=logodds)
        self._ck(self._lib.ipp_step_device(self._h, action_ids_ptr or None, poses_ptr or None, noise_ptr or None,
                                           noise_stride or self.max_measurements, reward_ptr or None, None, fl))

    def predict_device(self, n_jobs: int, action_ids_ptr: int = 0, poses_ptr: int = 0, env_index_ptr: int = 0, prev_ptr: int = 0,
                       reward_ptr: int = 0, commit=True, reward_mode=capi.REWARD_TRACE, adaptive=False) -> None:
        fl = self._flags(reward_mode, adaptive, commit=commit)
        self._ck(self._lib.ipp_predict_device(self._h, n_jobs, env_index_ptr or None, action_ids_ptr or None, poses_ptr or None,
                                              prev_ptr or None, reward_ptr or None, fl))
