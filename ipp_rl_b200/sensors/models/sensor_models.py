"""AltitudeSensorModel (reference sensors/models/sensor_models.py:13-85).

Host-side scalar helpers only: inside the engine the same quantities are evaluated by the fused
kernel (quad_math.cuh) from the per-altitude LUT.  ``measurement_model_matrix`` is provided for
callers that want the dense H of a footprint (tests, visualisation); the engine never builds it.
"""
import math
from typing import Tuple

import numpy as np

from . import SensorModel


class AltitudeSensorModel(SensorModel):
    def __init__(self, coeff_a: float, coeff_b: float):
        super().__init__()
        self.coeff_a = coeff_a  # noise level reached at high altitude
        self.coeff_b = coeff_b  # how fast it is reached

    def get_noise_variance(self, position: np.array) -> float:
        """sigma2(h) = a * (1 - exp(-b h))  (reference :27-30)"""
        return self.coeff_a * (1 - np.exp(-self.coeff_b * position[2]))

    def measurement_variance_matrix(self, position: np.array, num_measurements: int, resolution_factor: float) -> np.array:
        """R = rf^3 * sigma2(h) * I  (reference :32-36)"""
        return resolution_factor ** 3 * self.get_noise_variance(position) * np.identity(num_measurements)

    @staticmethod
    def measurement_blocks(field_of_view_indices: Tuple, resolution_factor: int):
        """Blocks of the measurement model in measurement order: (rows, cols, weight) with inclusive-exclusive
        cell ranges.  Block i sits at (i // nbx, i % nbx); a block with fewer than rf^2 cells weighs 1/rf
        instead of 1/rf^2 (reference :54-81)."""
        xl, xr, yu, yd = field_of_view_indices
        rf = int(resolution_factor)
        nx, ny = xr - xl + 1, yd - yu + 1
        nbx, nby = math.ceil(nx / rf), math.ceil(ny / rf)
        for i in range(nbx * nby):
            by, bx = divmod(i, nbx)
            r0, c0 = by * rf, bx * rf
            r1, c1 = min(r0 + rf, ny), min(c0 + rf, nx)
            full = (r1 - r0) * (c1 - c0) == rf * rf
            yield (yu + r0, yu + r1), (xl + c0, xl + c1), (1.0 / rf ** 2 if full else 1.0 / rf)

    def measurement_model_matrix(self, grid_map, field_of_view_indices: Tuple, num_measurements, resolution_factor: int) -> np.array:
        H = np.zeros((int(num_measurements), grid_map.num_grid_cells))
        for i, ((ra, rb), (ca, cb), w) in enumerate(self.measurement_blocks(field_of_view_indices, resolution_factor)):
            rows, cols = np.mgrid[ra:rb, ca:cb]
            H[i, (grid_map.x_dim * rows + cols).ravel()] = w
        return H
