"""GPU: the reference-shaped B=1 surface (GridMap / factories / Mapping / simulate_prediction_step /
greedy_search) on the CUDA engine.  These tests read like a reference experiment
(experiments/experiments.py:154-168, planning/greedy_mission.py:73-110)."""
import copy

import numpy as np
import pytest

from oracle import ipp_oracle as orc
from tests._util import golden, make_params, oracle_cfg, params_from_json

pytestmark = pytest.mark.gpu


def build(params, seed):
    """The reference's construction sequence, with this package's classes."""
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
    mapping = Mapping(grid_map, sensor)
    return grid_map, sensor, sensor_simulation, mapping


@pytest.mark.parametrize("name", ["ex10", "g24", "ns30x20"])
def test_reference_call_sequence_on_golden_episode(name):
    """take_measurement -> update_grid_map(predict_only) -> simulate_prediction_step, exactly the calls
    tests/golden/make_golden.py issued against the real reference, compared with its outputs."""
    from ipp_rl_b200.backend import drop_backend
    from ipp_rl_b200.planning.common.optimization import simulate_prediction_step

    g = golden("golden_episodes_T1.npz")
    params = params_from_json(g[f"ep_{name}_cfg"])
    params["sensor"].update(type="rgb_camera", encoding="rgb8", simulation={"type": "gaussian_random_field", "cluster_radius": 5})
    params["sensor"]["model"]["type"] = "altitude_dependent"
    params["mapping"] = make_params(4, 4, 1, 8, 14, 6)["mapping"]
    gm, sensor, sim, mapping = build(params, 1)
    sim.ground_truth_map = g[f"ep_{name}_gt"]
    uav = params["experiment"]["uav"]
    thr = params["experiment"]["scenario"]["value_threshold"]
    mean, var = g[f"ep_{name}_mean0"], g[f"ep_{name}_var0"]
    try:
        for t in range(min(12, len(g[f"ep_{name}_action"]))):
            a, prev = g[f"ep_{name}_action"][t], g[f"ep_{name}_prev"][t]
            np.random.seed(9000 + 17 * t)  # the seed the golden generator used for this step's noise
            z = sensor.take_measurement(a, verbose=False)
            m = int(np.prod(g[f"ep_{name}_zshape"][t]))
            assert z.shape == tuple(g[f"ep_{name}_zshape"][t])
            assert np.max(np.abs(z.ravel() - g[f"ep_{name}_z"][t][:m])) <= 1e-5
            P0 = np.diag(var.ravel())  # dense, as a reference caller would pass it
            gm.mean = mean.copy()
            r_plain, _, P1c = simulate_prediction_step(P0, prev, a, mapping, uav, None)
            adaptive_info = {"mean": mean.copy(), "value_threshold": thr, "interval_factor": 0.25}
            r_adapt, _, _ = simulate_prediction_step(P0, prev, a, mapping, uav, adaptive_info)
            x1, P1 = mapping.update_grid_map(a, z, cov_only=False, predict_only=True, current_cov_matrix=P0)
            assert np.max(np.abs(np.diag(P1) - np.diag(P1c))) == 0
            assert np.max(np.abs(x1 - g[f"ep_{name}_mean"][t])) <= 1e-5
            assert np.max(np.abs(np.diag(P1).reshape(var.shape) - g[f"ep_{name}_var"][t])) <= 1e-5
            assert abs(r_plain - g[f"ep_{name}_reward"][t]) <= 1e-5 * max(1.0, abs(g[f"ep_{name}_reward"][t]))
            margin = np.abs(mean + 0.25 * var - thr)
            if margin.min() > 1e-5:
                assert abs(r_adapt - g[f"ep_{name}_reward_adaptive"][t]) <= 1e-5 * max(1.0, abs(g[f"ep_{name}_reward_adaptive"][t]))
            mean, var = g[f"ep_{name}_mean"][t], g[f"ep_{name}_var"][t]
    finally:
        drop_backend(gm)


def test_greedy_mission_loop_against_oracle():
    """GreedyMission.execute's loop (reference planning/greedy_mission.py:73-110) with the engine-backed
    greedy_search (all candidates in one launch) against an oracle-driven twin."""
    from ipp_rl_b200.backend import drop_backend
    from ipp_rl_b200.planning.common.actions import action_costs
    from ipp_rl_b200.planning.common.optimization import greedy_search
    from ipp_rl_b200.planning.evaluation_metrics import map_uncertainty, root_mean_squared_error

    params = make_params(10, 10, 4, 8, 14, 6, thr=0.4, kappa=0.0)
    params["mapping"]["fit_gaussian_process"] = True
    gm, sensor, sim, mapping = build(params, 0)
    cfg = oracle_cfg(params)
    uav = params["experiment"]["uav"]
    o_gt = np.array(sim.ground_truth_map, dtype=np.float32).astype(np.float64)
    o_mean, o_var = gm.mean.copy(), gm.var.copy()
    assert np.allclose(o_var, 1.82) and np.allclose(o_mean, 0.5)
    previous_action, budget = np.array([2.0, 2.0, 14.0]), 60.0
    tbl = orc.enumerate_actions(cfg)
    steps = 0
    try:
        while budget >= 0 and steps < 12:
            adaptive_info = {"mean": gm.mean, "value_threshold": 0.4, "interval_factor": 0.0}
            wps = greedy_search(previous_action, budget, gm.cov_matrix, 1, mapping, 8, 14, 6, uav, adaptive_info=adaptive_info)
            if len(wps) == 0:
                break
            wp = np.array(wps[0])
            # oracle: rewards of every affordable action from the same state
            cand = [a for a in tbl if 0 < orc.action_costs(a, previous_action, cfg.uav) <= budget]
            ro = np.array([orc.simulate_prediction_step(cfg, o_var, previous_action, a, mean=o_mean, adaptive=True)[0] for a in cand])
            r_wp = orc.simulate_prediction_step(cfg, o_var, previous_action, wp, mean=o_mean, adaptive=True)[0]
            assert r_wp >= ro.max() * (1 - 1e-5), "greedy choice must be an (almost exact) arg-max of the oracle rewards"
            np.random.seed(100 + steps)
            z = sensor.take_measurement(wp, verbose=False)
            mapping.update_grid_map(wp, z)
            np.random.seed(100 + steps)
            eps = np.random.standard_normal(z.shape)
            _, o_mean, o_var, oz = orc.full_step(cfg, o_gt, o_mean, o_var, previous_action, wp, eps)
            assert np.max(np.abs(oz - z)) <= 1e-5
            assert np.max(np.abs(gm.mean - o_mean)) <= 1e-5 and np.max(np.abs(gm.var - o_var)) <= 1e-5
            o_mean, o_var = gm.mean.copy(), gm.var.copy()  # stay on the engine's fp32 trajectory
            budget -= action_costs(wp, previous_action, uav)
            previous_action = wp
            steps += 1
            assert map_uncertainty(gm.cov_matrix) == pytest.approx(o_var.sum(), rel=1e-6)
            assert np.isfinite(root_mean_squared_error(sim.ground_truth_map, gm.mean))
        assert steps >= 5
        assert np.trace(gm.cov_matrix) < 0.8 * 182.0  # the mission reduced the map uncertainty
    finally:
        drop_backend(gm)


def test_batched_twin_matches_b1_facade():
    """The batched engine and the B=1 facade are the same code path: one env stepped through both."""
    from ipp_rl_b200 import BatchedEngine, EngineConfig
    from ipp_rl_b200.backend import drop_backend

    params = make_params(24, 24, 1.0, 8, 20, 6)
    gm, sensor, sim, mapping = build(params, 3)
    eng = BatchedEngine(EngineConfig.from_params(params, batch=4, layout=1))
    try:
        eng.reset(0.5, 1.0)
        eng.set_state(np.broadcast_to(gm.mean, (4, 24, 24)), np.broadcast_to(gm.var, (4, 24, 24)))
        eng.set_ground_truth(np.broadcast_to(sim.ground_truth_map.astype(np.float32), (4, 24, 24)))
        rng = np.random.RandomState(0)
        for t in range(4):
            pose = np.array([rng.uniform(0, 24), rng.uniform(0, 24), [8.0, 14.0, 20.0][t % 3]])
            np.random.seed(t)
            z = sensor.take_measurement(pose, verbose=False)
            mapping.update_grid_map(pose, z)
            np.random.seed(t)
            eps = np.zeros((4, max(eng.max_measurements, z.size)), np.float32)
            eps[:, : z.size] = np.random.standard_normal(z.shape).ravel()
            eng.step(np.tile(pose, (4, 1)), noise=eps)
            m, v = eng.get_state()
            assert np.array_equal(m[0].astype(np.float64), gm.mean) and np.array_equal(v[2].astype(np.float64), gm.var)
    finally:
        eng.close()
        drop_backend(gm)
