"""Camera / RGBCamera (reference sensors/cameras.py:13-125).

``project_field_of_view`` is the host-side twin of the kernel's footprint decode (quad_math.cuh
``decode``): same float64 expression order, so both agree with the reference cell for cell.
"""
import logging
from typing import Dict, Tuple

import numpy as np

from . import Sensor

logger = logging.getLogger(__name__)

RESOLUTION_FACTOR_ALTITUDE = 10.0  # reference cameras.py:125


def footprint_cells(position, angle_x, angle_y, resolution, x_dim, y_dim) -> Tuple[int, int, int, int]:
    """(xl, xr, yu, yd): inclusive cell rectangle seen from pose [x, y, h] (reference :34-75).
    half-extent = floor(floor(2 h tan(angle/2) / res) / 2) cells around floor(pos / res), clipped."""
    h = position[2]
    span = np.array([2 * h * np.tan(0.5 * np.radians(angle_x)), 2 * h * np.tan(0.5 * np.radians(angle_y))])
    half = np.floor(0.5 * np.floor(span / resolution))
    centre = np.floor(np.asarray(position[:2], dtype=np.float64) / resolution)
    lo, hi = centre - half, centre + half
    xl, xr = np.clip([lo[0], hi[0]], 0, x_dim - 1)
    yu, yd = np.clip([lo[1], hi[1]], 0, y_dim - 1)
    return int(xl), int(xr), int(yu), int(yd)


class Camera(Sensor):
    def __init__(self, field_of_view: Dict, sensor_model, grid_map):
        super().__init__(sensor_model, grid_map)
        self.field_of_view = field_of_view

    @property
    def angle_x(self) -> float:
        return self.field_of_view["angle_x"]

    @property
    def angle_y(self) -> float:
        return self.field_of_view["angle_y"]

    def field_of_view_range(self, height: float) -> Tuple[float, float]:
        """ground-plane extent [m] of the FoV from `height` [m]"""
        return 2 * height * np.tan(0.5 * np.radians(self.angle_x)), 2 * height * np.tan(0.5 * np.radians(self.angle_y))

    def project_field_of_view(self, position: np.array) -> Tuple[int, int, int, int]:
        g = self.grid_map
        return footprint_cells(position, self.angle_x, self.angle_y, g.resolution, g.x_dim, g.y_dim)

    def take_measurement(self, position: np.array, verbose: bool = True) -> np.array:
        pass

    def process_measurement(self, image: np.array) -> np.array:
        pass

    def get_resolution_factor(self, position: np.array) -> float:
        pass


class RGBCamera(Camera):
    def __init__(self, field_of_view: Dict, sensor_model, grid_map, encoding: str = "rgb8"):
        super().__init__(field_of_view, sensor_model, grid_map)
        self.encoding = encoding

    def take_measurement(self, position: np.array, verbose: bool = True) -> np.array:
        """simulated measurement if a simulation is attached, else a random RGB image (reference :108-116)"""
        if verbose:
            logger.info(f"Take measurement at point: {position}")
        if self.sensor_simulation is None:
            return (np.random.random((self.grid_map.x_dim, self.grid_map.y_dim, 3)) * 255).astype(int)
        return self.sensor_simulation.take_measurement(position)

    def process_measurement(self, image: np.array) -> np.array:
        return image

    def get_resolution_factor(self, position: np.array) -> float:
        return 2 if position[2] > RESOLUTION_FACTOR_ALTITUDE else 1
