"""GPU: the mission surface on the engine against the REAL reference's runs (tests/golden/golden_missions.npz).

* C1 (BASELINE.json configs[0]): the reference's GreedyMission on config/example.yaml, covariance re-diagonalised after
  every update (the engine's belief model), reproduced waypoint for waypoint by this package's GreedyMission created through
  MissionFactory, with its six metric histories (device eval kernel) against the reference's.
* Mapping.init_priors: the non-GP and the shuffle_prior_cov branches (mapping/mappings.py:217-261).
* static Mapping.kalman_filter_update (mapping/mappings.py:155-215) on the reference's H / R.
* the deploy-time MCTSZeroMission loop runs and reduces the map uncertainty.
"""
import copy
import json

import numpy as np
import pytest

from tests._util import golden

pytestmark = pytest.mark.gpu


def build(params, seed, shuffle=False):
    from ipp_rl_b200.mapping.grid_maps import GridMap
    from ipp_rl_b200.mapping.mappings import Mapping
    from ipp_rl_b200.sensors.models.sensor_model_factories import SensorModelFactory
    from ipp_rl_b200.sensors.sensor_factories import SensorFactory
    from ipp_rl_b200.simulations.simulation_factories import SimulationFactory

    np.random.seed(seed)
    grid_map = GridMap(params)
    sensor_model = SensorModelFactory(params).create_sensor_model()
    sensor = SensorFactory(params, sensor_model, grid_map).create_sensor()
    sensor_simulation = SimulationFactory(params, sensor).create_sensor_simulation()
    sensor.set_sensor_simulation(sensor_simulation)
    return grid_map, sensor, sensor_simulation, Mapping(grid_map, sensor, shuffle_prior_cov=shuffle)


def _mission_params(params, kind, **extra):
    p = copy.deepcopy(params)
    m = dict(type=kind, config_name="standard")
    m.update(extra)
    m.update(p["experiment"]["constraints"])
    m.update(p["experiment"]["scenario"])
    p["mission"] = m
    return p


def test_greedy_mission_reproduces_the_reference_trace():
    from ipp_rl_b200.backend import drop_backend
    from ipp_rl_b200.planning.greedy_mission import GreedyMission
    from ipp_rl_b200.planning.mission_factories import MissionFactory

    g = golden("golden_missions.npz")
    params = json.loads(str(g["greedy_cfg"]))
    gm, sensor, sim, mapping = build(params, 5)
    try:
        sim.ground_truth_map = g["greedy_gt"]
        assert np.allclose(gm.var, 1.82)
        mission = MissionFactory(_mission_params(params, "greedy", num_waypoints=100), mapping, use_effective_mission_time=False).create_mission()
        assert isinstance(mission, GreedyMission) and mission.mission_label == "Greedy (standard)"
        np.random.seed(0)  # the seed the reference mission ran under: same measurement noise
        mission.execute()
        ref_wp = g["greedy_waypoints"]
        assert mission.waypoints.shape == ref_wp.shape, (len(mission.waypoints), len(ref_wp))
        assert np.array_equal(mission.waypoints, ref_wp)
        for mine, name, tol in ((mission.root_mean_squared_errors, "rmse", 2e-5), (mission.weighted_root_mean_squared_errors, "wrmse", 2e-5),
                                (mission.mean_log_losses, "mll", 2e-5), (mission.weighted_mean_log_losses, "wmll", 2e-5),
                                (mission.map_uncertainties, "unc", 2e-5), (mission.map_uncertainty_differences, "unc_diff", 2e-4)):
            ref = g[f"greedy_{name}"]
            assert len(mine) == len(ref) == len(ref_wp) + 1
            assert np.allclose(mine, ref, rtol=tol, atol=tol, equal_nan=True), (name, np.nanmax(np.abs(np.array(mine) - ref)))
        assert np.allclose(mission.flight_times, g["greedy_flight_times"], rtol=1e-12)
        assert np.max(np.abs(gm.mean - g["greedy_mean"])) <= 1e-5 and np.max(np.abs(gm.var - g["greedy_var"])) <= 1e-5
    finally:
        drop_backend(gm)


@pytest.mark.parametrize("name,gp,shuffle", [("nongp", False, False), ("nongp_shuffle", False, True), ("gp_shuffle", True, True)])
def test_prior_branches_match_the_reference(name, gp, shuffle):
    from ipp_rl_b200.backend import drop_backend
    from ipp_rl_b200.mapping.mappings import Mapping

    g = golden("golden_missions.npz")
    base = json.loads(str(g["greedy_cfg"]))
    base["mapping"]["fit_gaussian_process"] = gp
    gm, sensor, sim, _ = build(base, 3)
    try:
        np.random.seed(int(g[f"prior_{name}_seed"]))
        mapping = Mapping(gm, sensor, shuffle_prior_cov=shuffle)
        assert np.allclose(mapping.grid_map.cov_matrix.var, g[f"prior_{name}_diag"], rtol=1e-12, atol=0)
        assert np.array_equal(mapping.grid_map.mean, g[f"prior_{name}_mean"])
        m, v = __import__("ipp_rl_b200").backend.get_backend(gm).read_real()  # the prior reached the device (fp32)
        assert np.allclose(v.ravel(), g[f"prior_{name}_diag"], rtol=1e-6)
    finally:
        drop_backend(gm)


def test_static_kalman_filter_update_matches_the_reference():
    from ipp_rl_b200.mapping.grid_maps import DiagonalCovariance
    from ipp_rl_b200.mapping.mappings import Mapping

    g = golden("golden_missions.npz")
    for k in range(int(g["kf_n"])):
        var, mean, H, R, z = (g[f"kf_{k}_{n}"] for n in ("var", "mean", "H", "R", "z"))
        for P in (np.diag(var), DiagonalCovariance(var), var):
            x1, P1 = Mapping.kalman_filter_update(P, H, R, grid_mean=mean, observation=z, cov_only=False)
            assert np.max(np.abs(x1 - g[f"kf_{k}_x1"])) <= 1e-12
            assert np.max(np.abs(np.diag(P1) - g[f"kf_{k}_diagP1"])) <= 1e-12
        x0, P2 = Mapping.kalman_filter_update(np.diag(var), H, R, cov_only=True)
        assert x0 is None and np.max(np.abs(P2.var - g[f"kf_{k}_diagP1"])) <= 1e-12
    dense = np.diag(var).copy()
    dense[0, 1] = dense[1, 0] = 0.1
    with pytest.raises(ValueError):
        Mapping.kalman_filter_update(dense, H, R, cov_only=True)
    Hbad = H.copy()
    Hbad[0, np.flatnonzero(H[1])[0]] = 0.25
    with pytest.raises(ValueError):
        Mapping.kalman_filter_update(np.diag(var), Hbad, R, cov_only=True)


def test_mcts_zero_deploy_mission_runs_on_the_engine():
    from ipp_rl_b200.backend import drop_backend
    from ipp_rl_b200.planning.mcts_zero.mcts_zero_mission import MCTSZeroMission

    g = golden("golden_missions.npz")
    params = json.loads(str(g["greedy_cfg"]))
    gm, sensor, sim, mapping = build(params, 5)
    hyper = dict(gamma=1, puct_init=15, puct_base=10000, forced_playout_factor=2, num_mcts_simulations=24, max_valid_action_distance=11.5,
                 max_episode_steps=40, dirichlet_alpha=0.3, dirichlet_eps=0.25, num_workers=4)
    try:
        sim.ground_truth_map = g["greedy_gt"]
        mission = MCTSZeroMission(mapping, params["experiment"]["uav"], hyper, dist_to_boundaries=3, min_altitude=8, max_altitude=14,
                                  episode_horizon=3, altitude_spacing=6, budget=60, adaptive=True, value_threshold=0.4, interval_factor=0)
        assert mission.actions_np.shape == (200, 3) and mission.meta_data["num_grid_cells"] == 100
        msk = mission.get_next_actions_mask(np.array([2.0, 2.0, 14.0]), 60.0)
        assert msk.sum() > 0 and not msk[np.argmin(np.linalg.norm(mission.actions_np - [2, 2, 14], axis=1))]
        np.random.seed(1)
        mission.execute()
        assert len(mission.waypoints) >= 4
        d = np.linalg.norm(np.diff(np.vstack([[2, 2, 14], mission.waypoints]), axis=0), axis=1)
        assert np.all(d > 0) and np.all(d < 11.5)
        assert mission.map_uncertainties[-1] < mission.map_uncertainties[0]
        assert len(mission.root_mean_squared_errors) == len(mission.waypoints) + 1
    finally:
        drop_backend(gm)
