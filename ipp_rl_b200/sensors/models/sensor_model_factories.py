"""SensorModelFactory (reference sensors/models/sensor_model_factories.py:11-57)."""
from typing import Dict

from ..._config import require, require_member
from ...constants import REQUIRED_KEYS, SENSOR_MODELS, SensorModelType
from . import SensorModel
from .sensor_models import AltitudeSensorModel

_BUILDERS = {SensorModelType.ALTITUDE_DEPENDENT: AltitudeSensorModel}


class SensorModelFactory:
    def __init__(self, params: Dict):
        self.params = params
        self.model_params = self.get_model_params()

    @property
    def sensor_model(self) -> str:
        return require(self.params, ("sensor", "model", "type"), "sensor model type")

    def get_model_params(self) -> Dict:
        require_member(self.sensor_model, SENSOR_MODELS, "sensor models")
        return {k: require(self.params, ("sensor", "model", k), f"'{k}' parameter for sensor model '{self.sensor_model}'")
                for k in REQUIRED_KEYS[("model", self.sensor_model)]}

    def create_sensor_model(self) -> SensorModel:
        require_member(self.sensor_model, SENSOR_MODELS, "sensor models")
        return _BUILDERS[self.sensor_model](**self.model_params)
