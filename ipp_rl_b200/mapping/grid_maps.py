"""GridMap — belief container of the drop-in surface (reference mapping/grid_maps.py:7-54).

Same constructor and attributes: ``params``, ``mean`` (y_dim, x_dim), ``cov_matrix``, ``x_dim``,
``y_dim``, ``resolution``, ``num_grid_cells``.  The engine keeps the DIAGONAL of the covariance
(``var``); ``cov_matrix`` is a ``DiagonalCovariance`` that behaves like the reference's (N, N) array for
the operations its callers perform (``np.diag``, ``np.trace``, indexing, ``np.asarray``) and
materialises the dense matrix only when something asks for it.
"""
from typing import Dict

import numpy as np

from .._config import require


class DiagonalCovariance:
    """diag(var) with an ndarray-like face.  ``np.asarray(c)`` gives the dense (N, N) float64 matrix
    (what reference code such as planning/common/rewards.py:23-24 receives); ``c.var`` is the fast path."""

    __array_priority__ = 100.0

    def __init__(self, var: np.ndarray):
        self.var = np.ascontiguousarray(var, dtype=np.float64).ravel()

    @property
    def shape(self):
        n = self.var.size
        return (n, n)

    ndim = 2
    dtype = np.dtype(np.float64)

    def __array__(self, dtype=None, copy=None):
        dense = np.diag(self.var)
        return dense if dtype is None else dense.astype(dtype, copy=False)

    def diagonal(self, *a, **k):
        return self.var.copy()

    def trace(self, *a, **k):
        return float(self.var.sum())

    def copy(self):
        return DiagonalCovariance(self.var.copy())

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def __len__(self):
        return self.var.size


class VarianceMap(np.ndarray):
    """(y_dim, x_dim) per-cell variances — what ``GridMap.var`` returns.  The type tells ``covariance_diagonal`` that a
    SQUARE map is a variance map and not a dense (N, N) covariance (the two are indistinguishable by shape)."""

    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)


def covariance_diagonal(cov, num_cells: int = None) -> np.ndarray:
    """Diagonal of a covariance given as DiagonalCovariance, dense (N, N) array, VarianceMap, or (N,) variances.

    A plain square 2-D array is a dense covariance, as in the reference, unless ``num_cells`` says its SIZE is the number
    of grid cells (then it is a (Y, X) variance map); non-square 2-D arrays are variance maps."""
    if isinstance(cov, DiagonalCovariance):
        return cov.var
    if isinstance(cov, VarianceMap):
        return np.asarray(cov).ravel()
    a = np.asarray(cov, dtype=np.float64)
    if a.ndim == 2 and a.shape[0] == a.shape[1] and a.shape[0] > 1:
        if num_cells is not None and a.size == num_cells and a.shape[0] != num_cells:
            return a.ravel()
        return np.ascontiguousarray(np.diag(a))
    return a.ravel()


class GridMap:
    def __init__(self, params: Dict):
        self.params = params
        self.mean = None
        self.cov_matrix = None

    @property
    def x_dim(self) -> int:
        """map x-dimension in cells (environment.x_dim)"""
        return require(self.params, ("environment", "x_dim"))

    @property
    def y_dim(self) -> int:
        """map y-dimension in cells (environment.y_dim)"""
        return require(self.params, ("environment", "y_dim"))

    @property
    def resolution(self):
        """grid resolution in m/cell (environment.resolution)"""
        return require(self.params, ("environment", "resolution"))

    @property
    def num_grid_cells(self) -> int:
        return self.x_dim * self.y_dim

    @property
    def var(self) -> np.ndarray:
        """(y_dim, x_dim) view of the covariance diagonal — the quantity the engine stores."""
        return VarianceMap(covariance_diagonal(self.cov_matrix, self.num_grid_cells).reshape(self.y_dim, self.x_dim))

    # CUDA handles do not survive pickling / fork: drop the device backend, the child re-creates it lazily
    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_b200_backend", None)
        return state
