"""String registries of the hot-path factories (same keys and values as the reference's
constants.py:56-100, so one YAML drives both)."""


class SensorType:
    RGB_CAMERA = "rgb_camera"


class SensorModelType:
    ALTITUDE_DEPENDENT = "altitude_dependent"


class SensorSimulationType:
    GAUSSIAN_RANDOM_FIELD = "gaussian_random_field"
    HOTSPOT_RANDOM_FIELD = "hotspot_random_field"
    SPLIT_RANDOM_FIELD = "split_random_field"
    TEMPERATURE_DATA_FIELD = "temperature_data_field"


SENSOR_TYPES = [SensorType.RGB_CAMERA]
SENSOR_MODELS = [SensorModelType.ALTITUDE_DEPENDENT]
SENSOR_SIMULATIONS = [
    SensorSimulationType.GAUSSIAN_RANDOM_FIELD,
    SensorSimulationType.HOTSPOT_RANDOM_FIELD,
    SensorSimulationType.SPLIT_RANDOM_FIELD,
    SensorSimulationType.TEMPERATURE_DATA_FIELD,
]

# required config keys per registry entry (reference constants.py SensorParams / SensorModelParams /
# SensorSimulationParams)
REQUIRED_KEYS = {
    ("sensor", SensorType.RGB_CAMERA): ["field_of_view", "encoding"],
    ("model", SensorModelType.ALTITUDE_DEPENDENT): ["coeff_a", "coeff_b"],
    ("simulation", SensorSimulationType.GAUSSIAN_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.HOTSPOT_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.SPLIT_RANDOM_FIELD): ["cluster_radius"],
    ("simulation", SensorSimulationType.TEMPERATURE_DATA_FIELD): ["filename"],
}
