"""SimulationFactory (reference simulations/simulation_factories.py:12-75)."""
from typing import Dict

from .._config import require, require_member
from ..constants import REQUIRED_KEYS, SENSOR_SIMULATIONS, SensorSimulationType
from . import Simulation
from .simulations import GaussianRandomField, HotspotRandomField, SplitRandomField, TemperatureDataField

_BUILDERS = {
    SensorSimulationType.GAUSSIAN_RANDOM_FIELD: GaussianRandomField,
    SensorSimulationType.HOTSPOT_RANDOM_FIELD: HotspotRandomField,
    SensorSimulationType.SPLIT_RANDOM_FIELD: SplitRandomField,
    SensorSimulationType.TEMPERATURE_DATA_FIELD: TemperatureDataField,
}


class SimulationFactory:
    def __init__(self, params: Dict, sensor):
        self.params = params
        self.sensor = sensor
        self.simulation_params = self.get_simulation_params()

    @property
    def sensor_simulation(self) -> str:
        return require(self.params, ("sensor", "simulation", "type"), "sensor simulation type")

    def get_simulation_params(self) -> Dict:
        require_member(self.sensor_simulation, SENSOR_SIMULATIONS, "sensor simulations")
        out = {k: require(self.params, ("sensor", "simulation", k), f"'{k}' parameter for sensor simulation '{self.sensor_simulation}'")
               for k in REQUIRED_KEYS[("simulation", self.sensor_simulation)]}
        out["sensor"] = self.sensor
        return out

    def create_sensor_simulation(self) -> Simulation:
        require_member(self.sensor_simulation, SENSOR_SIMULATIONS, "sensor simulations")
        return _BUILDERS[self.sensor_simulation](**self.simulation_params)
