"""Sensor noise models (reference sensors/models/__init__.py:4-9)."""
import numpy as np


class SensorModel:
    def get_noise_variance(self, position: np.array) -> float:
        raise NotImplementedError("Sensor has no noise variance function implemented")
