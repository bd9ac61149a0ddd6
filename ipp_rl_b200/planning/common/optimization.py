"""Rollout step and greedy search (reference planning/common/optimization.py:14-104) on the engine.

``simulate_prediction_step`` keeps the reference signature.  ``greedy_search`` evaluates ALL candidate
actions of a step in one kernel launch (ipp_predict, NO_COMMIT) instead of a 4-process pool.
"""
from typing import Dict, List, Tuple

import numpy as np

from ...backend import get_backend
from ...mapping.grid_maps import DiagonalCovariance, covariance_diagonal
from .actions import action_costs, get_actions
from .rewards import compute_adaptive_msk, compute_reward


def simulate_prediction_step(current_state, previous_action, action, mapping, uav_specifications: Dict = None,
                             adaptive_info: Dict = None) -> Tuple[float, np.array, DiagonalCovariance]:
    adaptive_msk = None
    if adaptive_info is not None:
        adaptive_msk = compute_adaptive_msk(adaptive_info["mean"], current_state, adaptive_info["value_threshold"],
                                            adaptive_info["interval_factor"])
    _, next_state = mapping.update_grid_map(action, cov_only=True, predict_only=True, current_cov_matrix=current_state)
    reward = compute_reward(current_state, next_state, previous_action, action, uav_specifications, adaptive_msk)
    return reward, action, next_state


def greedy_search(previous_action, remaining_budget, current_state, episode_horizon, mapping, min_altitude, max_altitude,
                  altitude_spacing, uav_specifications: Dict = None, adaptive_info: Dict = None) -> List:
    """`episode_horizon` waypoints, each the reward-maximising affordable action from the state so far."""
    backend = get_backend(mapping.grid_map)
    if (uav_specifications is None) != (backend.cfg.max_v is None):
        raise ValueError("uav_specifications must match the experiment.uav section the engine was configured with")
    waypoints = []
    var = covariance_diagonal(current_state, mapping.grid_map.num_grid_cells)
    for _ in range(episode_horizon):
        candidates = get_actions(previous_action, remaining_budget, mapping.grid_map, min_altitude, max_altitude, altitude_spacing,
                                 uav_specifications)
        if len(candidates) == 0:
            break
        kw = {}
        if adaptive_info is not None:
            kw = dict(mean=adaptive_info["mean"], adaptive=True, value_threshold=adaptive_info["value_threshold"],
                      interval_factor=adaptive_info["interval_factor"])
        rewards = backend.rewards_from(var, previous_action, np.asarray(candidates), **kw)
        best = candidates[int(np.argmax(rewards))]  # first maximum, like the reference's strict '>' scan
        _, var = backend.predict_from(var, best)
        remaining_budget -= action_costs(best, previous_action, uav_specifications)
        previous_action = best
        waypoints.append(best)
    return waypoints
