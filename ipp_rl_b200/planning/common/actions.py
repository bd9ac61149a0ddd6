"""Action set and motion cost (reference planning/common/actions.py:8-106), vectorised.

The discrete action table is what the engine's integer action ids index:
``id = level * N + x_dim * col + row`` -> pose ``[res*col + res/2, res*row + res/2, altitude[level]]``.
"""
from typing import Dict, List, Optional

import numpy as np


def compute_distance(action: np.array, previous_action: np.array) -> float:
    return float(np.linalg.norm(np.asarray(action, float) - np.asarray(previous_action, float), ord=2))


def compute_flight_times(actions: np.array, previous_action: np.array, uav_specifications: Dict = None) -> np.ndarray:
    """Trapezoidal velocity profile: accelerate at max_a up to max_v, cruise, brake (reference :19-41)."""
    v, a = uav_specifications["max_v"], uav_specifications["max_a"]
    dist = np.linalg.norm(np.atleast_2d(actions) - np.asarray(previous_action, float), ord=2, axis=1)
    d_acc = np.minimum(np.square(v) / (2 * a), 0.5 * dist)
    return (dist - 2 * d_acc) / v + 2 * np.sqrt(2 * d_acc / a)


def compute_flight_time(action: np.array, previous_action: np.array, uav_specifications: Dict = None) -> float:
    return float(compute_flight_times(np.asarray(action, float)[None, :], previous_action, uav_specifications)[0])


def action_costs(action: np.array, previous_action: np.array, uav_specifications: Dict = None) -> float:
    if uav_specifications is None:
        return compute_distance(action, previous_action)
    return compute_flight_time(action, previous_action, uav_specifications)


def altitude_levels(min_altitude: float, max_altitude: float, altitude_spacing: float) -> np.ndarray:
    return np.linspace(min_altitude, max_altitude, int((max_altitude - min_altitude) / altitude_spacing) + 1)


def flatten_grid_index(grid_map, index_2d: np.array) -> int:
    return int(grid_map.x_dim * index_2d[0] + index_2d[1])


def action_table(grid_map, min_altitude: float, max_altitude: float, altitude_spacing: float) -> np.ndarray:
    """(levels * N, 3) poses indexed by the engine's action id (see module docstring)."""
    res, X, Y = grid_map.resolution, grid_map.x_dim, grid_map.y_dim
    lv = altitude_levels(min_altitude, max_altitude, altitude_spacing)
    n = X * Y
    table = np.zeros((len(lv) * n, 3))
    cols, rows = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")  # id within a level = X*col + row
    ids = (X * cols + rows).ravel()
    keep = ids < n  # non-square grids: the reference's id formula collides / overflows; keep what fits
    for h, alt in enumerate(lv):
        table[h * n + ids[keep], 0] = res * cols.ravel()[keep] + 0.5 * res
        table[h * n + ids[keep], 1] = res * rows.ravel()[keep] + 0.5 * res
        table[h * n + ids[keep], 2] = alt
    return table


def enumerate_actions(grid_map, min_altitude: float, max_altitude: float, altitude_spacing: float) -> Dict[int, np.ndarray]:
    """id -> pose dict like the reference (:73-91)."""
    return {i: row for i, row in enumerate(action_table(grid_map, min_altitude, max_altitude, altitude_spacing))}


def action_dict_to_np_array(actions: Dict) -> np.array:
    out = np.zeros((len(actions), 3))
    for idx, action in actions.items():
        out[idx, :] = action
    return out


def get_actions(previous_action, remaining_budget, grid_map, min_altitude, max_altitude, altitude_spacing,
                uav_specifications: Optional[Dict] = None) -> List[np.ndarray]:
    """All (cell centre, altitude level) poses reachable with 0 < cost <= budget, in the reference's order
    (row-major cells, altitude innermost; :44-66)."""
    res, X, Y = grid_map.resolution, grid_map.x_dim, grid_map.y_dim
    # the reference's candidate set steps by the spacing (:53-60), its id table uses linspace (:74): they differ whenever
    # (max - min) is not a multiple of the spacing
    lv = min_altitude + altitude_spacing * np.arange(int((max_altitude - min_altitude) / altitude_spacing) + 1)
    rows, cols, ks = np.meshgrid(np.arange(Y), np.arange(X), np.arange(len(lv)), indexing="ij")
    poses = np.stack([res * cols.ravel() + 0.5 * res, res * rows.ravel() + 0.5 * res, lv[ks.ravel()]], axis=1)
    prev = np.asarray(previous_action, float)
    cost = np.linalg.norm(poses - prev, axis=1) if uav_specifications is None else compute_flight_times(poses, prev, uav_specifications)
    return list(poses[(cost > 0) & (cost <= remaining_budget)])


def out_of_bounds(waypoint, grid_map, min_altitude: float, max_altitude: float):
    in_x = 0 <= waypoint[1] <= grid_map.x_dim * grid_map.resolution
    in_y = 0 <= waypoint[0] <= grid_map.y_dim * grid_map.resolution
    return not (in_x and in_y and min_altitude <= waypoint[2] <= max_altitude)
