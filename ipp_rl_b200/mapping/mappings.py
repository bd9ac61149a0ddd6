"""Mapping — Bayesian belief update of the drop-in surface (reference mapping/mappings.py:15-261).

Same constructor, properties and ``update_grid_map`` signature as the reference, executed by the
fused CUDA step kernel in its per-cell ("re-diagonalised Kalman") form:

* the belief is ``mean (y_dim, x_dim)`` + the DIAGONAL of the covariance; ``grid_map.cov_matrix`` is a
  ``DiagonalCovariance`` (acts like the reference's (N, N) array for ``np.diag`` / ``np.trace`` / ``np.asarray``);
* a dense covariance passed in (``current_cov_matrix``) is reduced to its diagonal — exact when the
  caller's matrix is diagonal, which is what a planner gets from this class.
"""
import logging
from typing import Optional, Tuple

import numpy as np

from .._config import require
from ..backend import get_backend
from .grid_maps import DiagonalCovariance, GridMap, covariance_diagonal

logger = logging.getLogger(__name__)

_MAPPING_KEYS = ("signal_variance", "noise_variance", "length_scale", "nu", "fit_gaussian_process", "prior_cov_mean", "prior_cov_std")


class Mapping:
    def __init__(self, grid_map: GridMap, sensor, shuffle_prior_cov: bool = False):
        self.grid_map = grid_map
        self.sensor = sensor
        self.shuffle_prior_cov = shuffle_prior_cov
        self.init_priors()

    def __getattr__(self, name):
        # mapping.<key> config properties of the reference (:22-112): looked up (and validated) on access
        if name in _MAPPING_KEYS:
            return require(self.grid_map.params, ("mapping", name), f"mapping's '{name}'")
        raise AttributeError(name)

    # -- priors (reference :217-261, diagonal restriction) ---------------------------------------------
    def init_priors(self):
        """mean = 0.5 everywhere; variance = diagonal of the reference's prior covariance:
        GP mode: the Matern kernel matrix has ``signal_variance`` on its diagonal (scaled by U(0.8, 1.2) when
        ``shuffle_prior_cov``); otherwise the diagonal of ``A A^T / ||A||_F`` with ``A ~ N(prior_cov_mean, prior_cov_std)``."""
        g = self.grid_map
        n = g.num_grid_cells
        if self.fit_gaussian_process:
            signal_variance = self.signal_variance
            if self.shuffle_prior_cov:
                signal_variance = np.random.uniform(low=0.8 * self.signal_variance, high=1.2 * self.signal_variance)
                np.random.uniform(low=0.8 * self.length_scale, high=1.2 * self.length_scale)  # keeps the RNG stream aligned
            var = np.full(n, float(signal_variance))
        else:
            mu, sd = self.prior_cov_mean, self.prior_cov_std
            if self.shuffle_prior_cov:
                mu = np.random.uniform(low=0.1, high=self.prior_cov_mean)
                sd = mu
            a = np.random.normal(mu, sd, (n, n))
            var = np.einsum("ij,ij->i", a, a) / np.linalg.norm(a, ord="fro")
        g.mean = 0.5 * np.ones((g.y_dim, g.x_dim))
        g.cov_matrix = DiagonalCovariance(var)
        self._push()

    def _push(self):
        get_backend(self.grid_map).load_real(self.grid_map.mean, self.grid_map.var)

    def _pull(self):
        mean, var = get_backend(self.grid_map).read_real()
        self.grid_map.mean = mean
        self.grid_map.cov_matrix = DiagonalCovariance(var)

    # -- the update --------------------------------------------------------------------------------------
    def update_grid_map(
        self,
        measurement_position: np.array,
        measurement_data: np.array = None,
        cov_only: bool = False,
        predict_only: bool = False,
        current_cov_matrix: np.array = None,
    ) -> Optional[Tuple[Optional[np.ndarray], DiagonalCovariance]]:
        backend = get_backend(self.grid_map)
        g = self.grid_map
        if predict_only:
            var = covariance_diagonal(g.cov_matrix if current_cov_matrix is None else current_cov_matrix, g.num_grid_cells)
            if cov_only:
                _, var_n = backend.predict_from(var, measurement_position)
                return None, DiagonalCovariance(var_n)
            mean_n, var_n = backend.predict_from(var, measurement_position, mean=g.mean, z=np.asarray(measurement_data))
            return mean_n, DiagonalCovariance(var_n)
        # committing update of the real belief; honour state the caller assigned to the grid map directly
        if current_cov_matrix is not None:
            g.cov_matrix = DiagonalCovariance(covariance_diagonal(current_cov_matrix, g.num_grid_cells))
        self._push()
        if cov_only:
            _, var_n = backend.predict_from(g.var, measurement_position)
            g.cov_matrix = DiagonalCovariance(var_n)
            self._push()
            return None
        backend.update(measurement_position, np.asarray(measurement_data))
        self._pull()
        return None

    @staticmethod
    def kalman_filter_update(P, H, R, grid_mean=None, observation=None, cov_only: bool = False, device: int = 0):
        """Static Kalman update (reference :155-215) for a DIAGONAL covariance and a measurement model whose rows have
        disjoint supports and one weight per row — what ``sensor_model.measurement_model_matrix`` /
        ``measurement_variance_matrix`` build.  Runs on the device (``ipp_kalman_blocks``, fp64): per block
        ``S = w^2 sum v + R``, ``v' = v - (w v)^2 / S``, ``x' = x + (w v / S)(z - w sum x)``; the off-diagonals the dense update
        would create inside a block are dropped.  Returns ``(x' flattened | None, DiagonalCovariance)``
        like the reference returns ``(x, P)``.  A dense, non-diagonal ``P`` or overlapping rows of ``H`` raise ValueError."""
        import ctypes as C

        from .. import _capi as capi

        H = np.asarray(H, dtype=np.float64)
        if H.ndim != 2:
            raise ValueError("H must be (num_measurements, num_grid_cells)")
        m, n = H.shape
        if not isinstance(P, DiagonalCovariance):
            dense = np.asarray(P, dtype=np.float64)
            if dense.shape == (n, n) and np.count_nonzero(dense - np.diag(np.diag(dense))):
                logger.error("kalman_filter_update: dense covariances with off-diagonal entries are not supported by the per-cell engine")
                raise ValueError("covariance must be diagonal")
        var = np.array(covariance_diagonal(P, num_cells=n), dtype=np.float64, copy=True)
        if var.size != n:
            raise ValueError(f"covariance has {var.size} diagonal entries, H has {n} columns")
        Rm = np.asarray(R, dtype=np.float64)
        r_diag = np.ascontiguousarray(np.diag(Rm) if Rm.ndim == 2 else np.broadcast_to(Rm, (m,)))
        rows, cols = np.nonzero(H)
        if rows.size and np.bincount(cols, minlength=n).max() > 1:
            raise ValueError("rows of H overlap: not a block measurement model")
        counts = np.bincount(rows, minlength=m)
        row_ptr = np.zeros(m + 1, np.int32)
        np.cumsum(counts, out=row_ptr[1:])
        weight = np.zeros(m)
        if rows.size:
            vals = H[rows, cols]
            first = np.minimum(row_ptr[:-1], max(rows.size - 1, 0))
            weight = np.where(counts > 0, vals[first], 0.0)
            if np.any(vals != weight[rows]):
                raise ValueError("rows of H must carry one weight each")
        cols32 = np.ascontiguousarray(cols, dtype=np.int32)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        x = z = None
        if not cov_only:
            x = np.array(np.asarray(grid_mean, dtype=np.float64).flatten(order="C"), copy=True)
            z = np.ascontiguousarray(np.asarray(observation, dtype=np.float64).flatten(order="C"))
            if x.size != n or z.size != m:
                raise ValueError("grid_mean / observation do not match H")
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        lib = capi.load_library()
        rc = lib.ipp_kalman_blocks(int(device), n, m, ptr(row_ptr), ptr(cols32), ptr(weight), ptr(r_diag), ptr(z), ptr(var), ptr(x))
        if rc != capi.IPP_OK:
            raise capi.IppError(rc, "ipp_kalman_blocks failed")
        return x, DiagonalCovariance(var)
