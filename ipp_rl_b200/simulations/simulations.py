"""Scalar-field simulations (reference simulations/simulations.py:16-168).

``take_measurement`` — crop the footprint, INTER_AREA down-sample above 10 m, add altitude-dependent
Gaussian noise, clip — runs on the GPU (the measurement half of the fused step kernel).  The noise
comes from ``np.random.standard_normal`` exactly where the reference calls ``np.random.normal``, so a
seeded run sees the same measurements as the reference (to fp32).
"""
import logging
import os

import numpy as np

from ..backend import get_backend, measurement_shape
from . import Simulation, ground_truths

logger = logging.getLogger(__name__)


class ScalarFieldSimulation(Simulation):
    def __init__(self, sensor, cluster_radius: float = None):
        super().__init__(sensor)
        self.cluster_radius = cluster_radius

    def create_ground_truth_map(self) -> np.array:
        raise NotImplementedError("Scalar field simulation has no function implemented to create ground truth map")

    def take_measurement(self, position: np.array, verbose: bool = True) -> np.array:
        position = np.asarray(position, dtype=np.float64)
        fov = self.sensor.project_field_of_view(position)
        shape = measurement_shape(fov, self.sensor.get_resolution_factor(position))
        eps = np.random.standard_normal(shape)  # == np.random.normal(0, s, shape) / s under the same seed
        backend = get_backend(self.sensor.grid_map)
        backend.sync_ground_truth(self.ground_truth_map)
        return backend.measure(position, eps)


class GaussianRandomField(ScalarFieldSimulation):
    def __init__(self, sensor, cluster_radius: float):
        super().__init__(sensor, cluster_radius)
        self.ground_truth_map = self.create_ground_truth_map()

    def create_ground_truth_map(self) -> np.array:
        """random field with spectrum k^-cluster_radius, values in [0, 1], shape (y_dim, x_dim)"""
        g = self.sensor.grid_map
        return ground_truths.gaussian_random_field(lambda k: k ** (-self.cluster_radius), g.x_dim, g.y_dim)


def _window(center: int, radius, limit: int):
    return int(max(center - radius, 0)), int(min(center + radius, limit))


class HotspotRandomField(ScalarFieldSimulation):
    def __init__(self, sensor, cluster_radius: float):
        super().__init__(sensor, cluster_radius)
        self.ground_truth_map = self.create_ground_truth_map()

    def create_ground_truth_map(self) -> np.array:
        """two square hot spots of side 2*cluster_radius on a low background (reference :57-92); same RNG
        call order: high, low, centre (y, x), then candidate second centres until both axes are clear."""
        g, r = self.sensor.grid_map, self.cluster_radius
        high, low = np.random.uniform(0.7, 1), np.random.uniform(0.0, 0.3)
        field = np.full((g.y_dim, g.x_dim), low)
        cy, cx = np.random.randint(r, g.y_dim), np.random.randint(r, g.x_dim)
        (y0, y1), (x0, x1) = _window(cy, r, g.y_dim), _window(cx, r, g.x_dim)
        field[y0:y1, x0:x1] = high
        while True:
            ty, tx = np.random.randint(r, g.y_dim), np.random.randint(r, g.x_dim)
            if abs(ty - cy) <= r or abs(tx - cx) <= r:
                continue
            (y0, y1), (x0, x1) = _window(ty, r, g.y_dim), _window(tx, r, g.x_dim)
            field[y0:y1, x0:x1] = high
            return field


class SplitRandomField(ScalarFieldSimulation):
    def __init__(self, sensor, cluster_radius: float):
        super().__init__(sensor, cluster_radius)
        self.ground_truth_map = self.create_ground_truth_map()

    def create_ground_truth_map(self) -> np.array:
        """field split into a high and a low half along a random row or column (reference :102-125)"""
        g = self.sensor.grid_map
        high, low = np.random.uniform(0.65, 1), np.random.uniform(0.0, 0.35)
        first, second = (low, high) if np.random.rand() > 0.5 else (high, low)
        field = np.ones((g.y_dim, g.x_dim))
        if np.random.rand() > 0.5:
            cut = np.random.randint(np.ceil(g.y_dim * 0.33), np.ceil(g.y_dim * 0.66) + 1)
            field[:cut, :], field[cut:, :] = first, second
        else:
            cut = np.random.randint(np.floor(g.x_dim * 0.33), np.ceil(g.x_dim * 0.66) + 1)
            field[:, :cut], field[:, cut:] = first, second
        return field


class TemperatureDataField(ScalarFieldSimulation):
    """Ground truth from an RGBA temperature image (reference :128-168).  The data set is not shipped
    with the reference; the loader needs an image reader (imageio or cv2) and the file under DATASETS_DIR."""

    def __init__(self, sensor, filename: str):
        super().__init__(sensor)
        self.raw_data = self.load_raw_data(filename)
        self.ground_truth_map = self.create_ground_truth_map()

    @staticmethod
    def load_raw_data(filename: str) -> np.array:
        path = os.path.join(os.environ.get("DATASETS_DIR", "datasets"), filename)
        if not os.path.exists(path):
            logger.error(f"Cannot find temperature ground truth data! File {path} does not exist!")
            raise ValueError
        try:
            import imageio

            return np.asarray(imageio.imread(path))
        except ImportError:
            import cv2

            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            return img[..., [2, 1, 0] + ([3] if img.shape[-1] == 4 else [])]  # BGR(A) -> RGB(A)

    @staticmethod
    def rgba_to_temperature(rgba: np.array) -> np.array:
        return -1 * (rgba[:, :, 0] - rgba[:, :, 2])

    @staticmethod
    def normalize_temperature_map(t: np.array) -> np.array:
        lo, hi = np.min(t), np.max(t)
        return t / hi if lo == hi else (t - lo) / (hi - lo)

    def create_ground_truth_map(self) -> np.array:
        import cv2

        g = self.sensor.grid_map
        t = self.normalize_temperature_map(self.rgba_to_temperature(self.raw_data))
        small = cv2.resize(t, dsize=(g.y_dim, g.x_dim), interpolation=cv2.INTER_AREA)  # dsize order as in the reference
        return self.normalize_temperature_map(small)
