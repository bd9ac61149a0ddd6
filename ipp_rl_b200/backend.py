"""B = 1 bridge between the reference-shaped objects (GridMap / Sensor / Simulation / Mapping) and the
batched CUDA engine.

One ``SingleEnvBackend`` per GridMap, created lazily by whoever needs the device first (a Simulation
taking a measurement, or the Mapping).  It owns two single-env engines: ``real`` holds the belief
the Mapping commits to, ``scratch`` serves ``predict_only`` calls that start from an arbitrary
covariance (reference mapping/mappings.py:114-153, planning/common/optimization.py:14-30) and the
many-candidates greedy search.  CUDA handles are neither picklable nor fork-safe: pickling drops the
backend and the child re-creates it on first use (threading / processes).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

from .engine import BatchedEngine, EngineConfig

_ATTR = "_b200_backend"


def measurement_shape(fov: Tuple[int, int, int, int], rf: int) -> Tuple[int, int]:
    """Shape of the simulated measurement (reference simulations/sensor_manipulations.py:18-24): identity at
    rf == 1; at rf > 1 cv2's dsize=(ceil(ny/rf), ceil(nx/rf)) is (width, height), i.e. the array has
    ceil(nx/rf) rows and ceil(ny/rf) columns."""
    xl, xr, yu, yd = fov
    nx, ny = xr - xl + 1, yd - yu + 1
    if rf <= 1:
        return ny, nx
    return math.ceil(nx / rf), math.ceil(ny / rf)


class SingleEnvBackend:
    def __init__(self, grid_map, scratch_jobs: int = 1):
        self.grid_map = grid_map
        self.cfg = EngineConfig.from_params(grid_map.params, batch=1, layout=1)
        self._real: Optional[BatchedEngine] = None
        self._scratch: Optional[BatchedEngine] = None
        self._gt_token = None
        self._gt_token_scratch = None

    # -- engines --------------------------------------------------------------------------------
    @property
    def real(self) -> BatchedEngine:
        if self._real is None:
            self._real = BatchedEngine(self.cfg)
            self._real.reset(0.5, 1.0)
        return self._real

    @property
    def scratch(self) -> BatchedEngine:
        if self._scratch is None:
            self._scratch = BatchedEngine(self.cfg)
            self._scratch.reset(0.5, 1.0)
        return self._scratch

    def close(self) -> None:
        for e in (self._real, self._scratch):
            if e is not None:
                e.close()
        self._real = self._scratch = None

    # -- world ----------------------------------------------------------------------------------
    def sync_ground_truth(self, gt: np.ndarray) -> None:
        """Upload the simulation's ground-truth map when it changed (identity + content hash)."""
        g = np.ascontiguousarray(gt, dtype=np.float32)
        if g.shape != (self.cfg.y_dim, self.cfg.x_dim):
            raise ValueError(f"ground truth map must have shape ({self.cfg.y_dim}, {self.cfg.x_dim}), got {g.shape}")
        token = (id(gt), hash(g.tobytes()))
        if token != self._gt_token:
            self.real.set_ground_truth(g)
            self._gt_token = token

    # -- hot path pieces ---------------------------------------------------------------------------
    def measure(self, position, eps: np.ndarray) -> np.ndarray:
        """Noisy footprint measurement with the supplied standard normals (C order of the measurement)."""
        e = self.real
        noise = np.zeros((1, max(e.max_measurements, eps.size)), np.float32)
        noise[0, : eps.size] = np.asarray(eps, np.float32).ravel()
        z = e.measure(np.asarray(position, np.float64).reshape(1, 3), noise=noise)
        return z[0, : eps.size].astype(np.float64).reshape(eps.shape)

    def update(self, position, z: np.ndarray) -> None:
        e = self.real
        zz = np.zeros((1, max(e.max_measurements, z.size)), np.float32)
        zz[0, : z.size] = np.asarray(z, np.float32).ravel(order="C")
        e.update(np.asarray(position, np.float64).reshape(1, 3), zz)

    def load_real(self, mean: np.ndarray, var: np.ndarray) -> None:
        self.real.set_state(np.asarray(mean, np.float32)[None], np.asarray(var, np.float32).reshape(1, self.cfg.y_dim, self.cfg.x_dim))

    def read_real(self) -> Tuple[np.ndarray, np.ndarray]:
        m, v = self.real.get_state()
        return m[0].astype(np.float64), v[0].astype(np.float64)

    def predict_from(self, var: np.ndarray, position, mean: Optional[np.ndarray] = None, z: Optional[np.ndarray] = None):
        """Reference ``predict_only`` update from an arbitrary diagonal state: returns (mean' or None, var')."""
        s = self.scratch
        Y, X = self.cfg.y_dim, self.cfg.x_dim
        s.set_state(None if mean is None else np.asarray(mean, np.float32).reshape(1, Y, X), np.asarray(var, np.float32).reshape(1, Y, X))
        pos = np.asarray(position, np.float64).reshape(1, 3)
        if z is None:
            s.predict(pos, commit=True)
        else:
            zz = np.zeros((1, max(s.max_measurements, z.size)), np.float32)
            zz[0, : z.size] = np.asarray(z, np.float32).ravel(order="C")
            s.update(pos, zz, keep_prev=True)
        m, v = s.get_state()
        return (None if z is None else m[0].astype(np.float64)), v[0].astype(np.float64)

    def set_mask_params(self, value_threshold=None, interval_factor=None) -> None:
        """The adaptive-mask parameters are engine configuration (ipp_config.value_threshold / interval_factor): rebuild the
        engines, keeping their state, when a caller asks for different ones (None = keep)."""
        thr = self.cfg.value_threshold if value_threshold is None else float(value_threshold)
        kap = self.cfg.interval_factor if interval_factor is None else float(interval_factor)
        if thr == self.cfg.value_threshold and kap == self.cfg.interval_factor:
            return
        self.cfg.value_threshold, self.cfg.interval_factor = thr, kap
        if self._scratch is not None:
            self._scratch.close()
            self._scratch = None
        if self._real is not None:  # keep both engines on the same configuration
            m, v = self.read_real()
            gt = self._real.get_ground_truth()
            prev = self._real.get_prev_pose()
            self._real.close()
            self._real = None
            self.real.set_ground_truth(gt)
            self.load_real(m, v)
            self.real.set_prev_pose(prev)

    def rewards_from(self, var: np.ndarray, previous_action, actions: np.ndarray, mean: Optional[np.ndarray] = None,
                     adaptive: bool = False, value_threshold: float = None, interval_factor: float = None) -> np.ndarray:
        """Information-gain rewards of many candidate actions from ONE state in a single launch
        (greedy_search's Pool(4) loop, reference planning/common/optimization.py:82-98)."""
        if adaptive:
            self.set_mask_params(value_threshold, interval_factor)
        s = self.scratch
        Y, X = self.cfg.y_dim, self.cfg.x_dim
        s.set_state(None if mean is None else np.asarray(mean, np.float32).reshape(1, Y, X), np.asarray(var, np.float32).reshape(1, Y, X))
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(-1, 3)
        return s.predict(a, env_index=np.zeros(len(a), np.int32), prev_poses=np.asarray(previous_action, np.float64), commit=False,
                         adaptive=adaptive).astype(np.float64)


def get_backend(grid_map) -> SingleEnvBackend:
    b = getattr(grid_map, _ATTR, None)
    if b is None:
        b = SingleEnvBackend(grid_map)
        setattr(grid_map, _ATTR, b)
    return b


def drop_backend(grid_map) -> None:
    b = getattr(grid_map, _ATTR, None)
    if b is not None:
        b.close()
        setattr(grid_map, _ATTR, None)
