"""Information-gain reward (reference planning/common/rewards.py:8-39) on diagonal states."""
from typing import Dict, Union

import numpy as np

from ...mapping.grid_maps import covariance_diagonal
from .actions import action_costs


def compute_adaptive_msk(grid_mean: np.array, grid_covariance, value_threshold: float, interval_factor: float):
    """cells whose upper confidence value mean + k * variance reaches the threshold (variance, as in the reference)"""
    return np.asarray(grid_mean).flatten(order="C") + interval_factor * covariance_diagonal(grid_covariance, int(np.size(grid_mean))) >= value_threshold


def compute_reward(current_state, next_state, previous_action, action, uav_specifications: Dict = None, adaptive_msk=None) -> float:
    """trace reduction over the (masked) cells per unit cost + 1"""
    before, after = covariance_diagonal(current_state), covariance_diagonal(next_state)
    if adaptive_msk is not None:
        before, after = before[adaptive_msk], after[adaptive_msk]
    return (np.sum(before) - np.sum(after)) / (action_costs(action, previous_action, uav_specifications) + 1)


def scale_value_target(value: float) -> float:
    return np.sqrt(value + 1) - 1


def invert_scaled_value_target(value: Union[float, np.array]) -> Union[float, np.array]:
    return np.square(value) + 2 * value
