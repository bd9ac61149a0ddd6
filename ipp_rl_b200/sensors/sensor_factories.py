"""SensorFactory (reference sensors/sensor_factories.py:12-67)."""
from typing import Dict

from .._config import require, require_member
from ..constants import REQUIRED_KEYS, SENSOR_TYPES, SensorType
from . import Sensor, cameras

_BUILDERS = {SensorType.RGB_CAMERA: cameras.RGBCamera}


class SensorFactory:
    def __init__(self, params: Dict, sensor_model, grid_map):
        self.params = params
        self.sensor_model = sensor_model
        self.grid_map = grid_map
        self.sensor_params = self.get_sensor_params()

    @property
    def sensor_type(self) -> str:
        return require(self.params, ("sensor", "type"), "sensor type")

    def get_sensor_params(self) -> Dict:
        require_member(self.sensor_type, SENSOR_TYPES, "sensor types")
        out = {k: require(self.params, ("sensor", k), f"'{k}' parameter for sensor type '{self.sensor_type}'")
               for k in REQUIRED_KEYS[("sensor", self.sensor_type)]}
        out["sensor_model"] = self.sensor_model
        out["grid_map"] = self.grid_map
        return out

    def create_sensor(self) -> Sensor:
        require_member(self.sensor_type, SENSOR_TYPES, "sensor types")
        return _BUILDERS[self.sensor_type](**self.sensor_params)
