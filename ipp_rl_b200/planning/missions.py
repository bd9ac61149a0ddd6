"""Mission — the interface every planning mission shares (reference planning/missions.py:22-239, without the
matplotlib views).  Same constructor, attributes and metric histories; ``eval`` takes its numbers from the
device (``ipp_eval``: eval_kernel in csrc/ipp_engine.cu), one launch per executed step.
"""
from typing import Dict, Optional

import numpy as np

from ..backend import get_backend


class Mission:
    def __init__(self, mapping, uav_specifications: Dict, dist_to_boundaries: float = 10, min_altitude: float = 5,
                 max_altitude: float = 30, budget: float = 100, adaptive: bool = False, value_threshold: float = 0.5,
                 interval_factor: float = 2, config_name: str = "standard", use_effective_mission_time: bool = False):
        self.mapping = mapping
        self.uav_specifications = uav_specifications
        self.dist_to_boundaries = dist_to_boundaries
        self.min_altitude = min_altitude
        self.max_altitude = max_altitude
        self.budget = budget
        self.adaptive = adaptive
        self.value_threshold = value_threshold
        self.interval_factor = interval_factor
        self.waypoints = np.empty((0, 3))
        self.config_name = config_name
        self.use_effective_mission_time = use_effective_mission_time
        self.mission_type = None
        self.mission_name = None
        self.init_action = np.array([2, 2, 14])  # planning/missions.py:69

        self.root_mean_squared_errors = []
        self.weighted_root_mean_squared_errors = []
        self.mean_log_losses = []
        self.weighted_mean_log_losses = []
        self.map_uncertainties = []
        self.map_uncertainty_differences = []
        self.run_times = []
        self.flight_times = []

    def create_waypoints(self) -> np.array:
        raise NotImplementedError("Planning mission does not implement 'create_waypoints' function!")

    def execute(self):
        raise NotImplementedError("Planning mission does not implement 'execute' function!")

    def get_adaptive_info(self) -> Optional[Dict]:
        if not self.adaptive:
            return None
        return {"mean": self.mapping.grid_map.mean, "value_threshold": self.value_threshold, "interval_factor": self.interval_factor}

    def eval(self, run_time: float = None, flight_time: float = None):
        """Evaluation metrics of the current map estimate, appended to the histories (reference :176-203).  The adaptive
        variants restrict RMSE / tr(P) to the cells with ``ground truth >= value_threshold``."""
        backend = get_backend(self.mapping.grid_map)
        backend.sync_ground_truth(self.mapping.sensor.sensor_simulation.ground_truth_map)
        backend.set_mask_params(self.value_threshold, None)
        self.mapping._push()
        m = backend.real.eval()[0].astype(np.float64)
        self.root_mean_squared_errors.append(m[6] if self.adaptive else m[0])
        self.weighted_root_mean_squared_errors.append(m[1])
        self.mean_log_losses.append(m[2])
        self.weighted_mean_log_losses.append(m[3])
        self.map_uncertainties.append(m[7] if self.adaptive else m[4])
        if self.adaptive:
            self.map_uncertainty_differences.append(m[5])
        if run_time is not None:
            self.run_times.append(run_time)
        if flight_time is not None:
            self.flight_times.append(flight_time)

    @property
    def mission_label(self):
        return f"{self.mission_name} ({self.config_name})"
