#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REAL REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (dmar-bonn/ipp-rl @ 25dfb33) is imported in place, with empty stub modules for
plotting / messaging dependencies that are not on the hot path.  Every
array written here is an output of the reference's own functions:

* ``Camera.project_field_of_view`` / ``get_resolution_factor``          sensors/cameras.py:49-75,122-125
* ``AltitudeSensorModel.get_noise_variance`` / ``measurement_model_matrix``  sensors/models/sensor_models.py:27-85
* ``ScalarFieldSimulation.take_measurement``                            simulations/simulations.py:26-34
* ``Mapping.update_grid_map`` (dense KF) started from ``np.diag(var)``   mapping/mappings.py:114-215   (tier T1)
* static ``Mapping.kalman_filter_update`` on the FoV window              (tier T2, 200x200)
* ``simulate_prediction_step`` with/without ``adaptive_info``            planning/common/optimization.py:14-30
* ``enumerate_actions`` / ``action_costs``                              planning/common/actions.py
* ``planning/evaluation_metrics.py`` reductions
* a dense (unmodified, GP prior) run for the SURVEY Appendix-B known answers (tier T0).

While generating, the NumPy oracle (oracle/ipp_oracle.py) is checked against each output and the
worst deviations are printed and stored in ``golden_meta.json``.
"""
import json
import os
import sys
import types

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
for _n in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "cma", "telegram"]:
    sys.modules[_n] = types.ModuleType(_n)
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

import copy  # noqa: E402

import numpy as np  # noqa: E402
import yaml  # noqa: E402

from mapping.grid_maps import GridMap  # noqa: E402  (reference)
from mapping.mappings import Mapping  # noqa: E402
from planning import evaluation_metrics as ref_metrics  # noqa: E402
from planning.common import actions as ref_actions  # noqa: E402
from planning.common.optimization import simulate_prediction_step  # noqa: E402
from sensors.models.sensor_model_factories import SensorModelFactory  # noqa: E402
from sensors.sensor_factories import SensorFactory  # noqa: E402
from simulations.simulation_factories import SimulationFactory  # noqa: E402

from oracle import ipp_oracle as orc  # noqa: E402

BASE = yaml.safe_load(open(os.path.join(REF, "config/example.yaml")))
WORST = {}


def note(name, err):
    WORST[name] = max(WORST.get(name, 0.0), float(err))


def make_params(x_dim, y_dim, res, alt_min, alt_max, alt_step, gp=False, angle=(60, 60)):
    p = copy.deepcopy(BASE)
    p["environment"].update(x_dim=x_dim, y_dim=y_dim, resolution=res)
    p["sensor"]["field_of_view"].update(angle_x=angle[0], angle_y=angle[1])
    p["mapping"]["fit_gaussian_process"] = gp
    p["experiment"]["constraints"].update(min_altitude=alt_min, max_altitude=alt_max, altitude_spacing=alt_step)
    return p


def build(params, seed):
    np.random.seed(seed)
    gm = GridMap(params)
    model = SensorModelFactory(params).create_sensor_model()
    sensor = SensorFactory(params, model, gm).create_sensor()
    sim = SimulationFactory(params, sensor).create_sensor_simulation()
    sensor.set_sensor_simulation(sim)
    mapping = Mapping(gm, sensor)
    return gm, sensor, sim, mapping


def cfg_json(params):
    keep = {
        "environment": params["environment"],
        "sensor": {k: params["sensor"][k] for k in ("field_of_view", "model")},
        "experiment": {k: params["experiment"][k] for k in ("constraints", "scenario", "uav")},
    }
    return json.dumps(keep)


# ----------------------------------------------------------------------------------------------
def gen_footprints(out):
    """Footprint / rf / sigma2 / H-structure over pose sweeps incl. every border."""
    scen = [
        make_params(10, 10, 4, 8, 14, 6),
        make_params(24, 24, 1, 8, 20, 6),
        make_params(200, 200, 1, 8, 20, 6),
        make_params(30, 20, 2, 8, 26, 9),
        make_params(16, 16, 1, 5, 25, 5, angle=(60, 60)),
        make_params(400, 400, 1, 8, 20, 6),
    ]
    rng = np.random.RandomState(7)
    for si, p in enumerate(scen):
        gm, sensor, sim, _ = build(p, 100 + si) if p["environment"]["x_dim"] <= 30 else (None, None, None, None)
        if gm is None:  # big grids: skip Mapping (dense prior) — sensor only
            gm = GridMap(p)
            model = SensorModelFactory(p).create_sensor_model()
            sensor = SensorFactory(p, model, gm).create_sensor()
        cfg = orc.OracleConfig.from_params(p)
        X, Y, res = gm.x_dim, gm.y_dim, gm.resolution
        poses = []
        alts = list(orc.altitude_levels(cfg)) + [10.0, 10.000001, 3.3, 17.77]
        for _ in range(150):
            poses.append([rng.uniform(0, X * res), rng.uniform(0, Y * res), alts[rng.randint(len(alts))]])
        for cx in (0, 1, X // 2, X - 2, X - 1):
            for cy in (0, 1, Y // 2, Y - 2, Y - 1):
                for a in alts[:4]:
                    poses.append([res * cx + 0.5 * res, res * cy + 0.5 * res, a])
        poses = np.array(poses, dtype=np.float64)
        fov = np.array([sensor.project_field_of_view(q) for q in poses], dtype=np.int32)
        rf = np.array([sensor.get_resolution_factor(q) for q in poses], dtype=np.int32)
        s2 = np.array([sensor.sensor_model.get_noise_variance(q) for q in poses])
        Rd = np.array([sensor.sensor_model.measurement_variance_matrix(q, 1, r)[0, 0] for q, r in zip(poses, rf)])
        out[f"fp{si}_cfg"] = cfg_json(p)
        out[f"fp{si}_poses"] = poses
        out[f"fp{si}_fov"] = fov
        out[f"fp{si}_rf"] = rf
        out[f"fp{si}_sigma2"] = s2
        out[f"fp{si}_R"] = Rd
        for q, f, r, s, rr in zip(poses, fov, rf, s2, Rd):
            assert orc.project_field_of_view(cfg, q) == tuple(int(t) for t in f), (q, f)
            assert orc.resolution_factor(cfg, q) == r
            note("sigma2", abs(orc.noise_variance(cfg, q) - s))
            note("R", abs(orc.measurement_variance(cfg, q, int(r)) - rr))
        # H structure (dense reference matrix) on small grids for a subset
        if X * Y <= 600:
            hs = []
            for q, f, r in list(zip(poses, fov, rf))[::7]:
                m = orc.num_measurements(tuple(f), int(r))
                H = sensor.sensor_model.measurement_model_matrix(gm, tuple(int(t) for t in f), m, int(r))
                Ho = orc.measurement_model_matrix(cfg, tuple(int(t) for t in f), int(r))
                assert H.shape == Ho.shape and np.array_equal(H, Ho), (q, f, r)
                hs.append([H.shape[0], int(np.count_nonzero(H)), float(H.sum())])
            out[f"fp{si}_Hstats"] = np.array(hs)
    out["fp_count"] = len(scen)


# ----------------------------------------------------------------------------------------------
def gen_episodes(out):
    """Tier T1: multi-step episodes where every step is the reference's dense update started
    from np.diag(var) (re-diagonalised Kalman step) — mean, var, z, reward, adaptive reward."""
    scen = [
        ("ex10", make_params(10, 10, 4, 8, 14, 6), 24, "half"),
        ("g24", make_params(24, 24, 1, 8, 20, 6), 40, "rand"),
        ("ns30x20", make_params(30, 20, 2, 8, 26, 9), 30, "rand"),
        ("g16", make_params(16, 16, 1, 5, 25, 5), 30, "rand"),
    ]
    rng = np.random.RandomState(11)
    for name, p, T, mean_init in scen:
        gm, sensor, sim, mapping = build(p, 500 + T)
        cfg = orc.OracleConfig.from_params(p)
        X, Y, res = gm.x_dim, gm.y_dim, gm.resolution
        uav = p["experiment"]["uav"]
        if X != Y:
            # reference quirk (Appendix C #7): GaussianRandomField swaps its dims, so on a non-square
            # grid its GT has shape (X, Y) and FoV slices come back empty.  Install a (Y, X) field
            # from the reference's own generator called with the un-swapped dims.
            from simulations import ground_truths as ref_gt

            sim.ground_truth_map = ref_gt.gaussian_random_field(lambda k: k ** (-5.0), X, Y)
            assert sim.ground_truth_map.shape == (Y, X)
        gt = np.array(sim.ground_truth_map, copy=True)
        var = rng.uniform(0.1, 2.0, size=(Y, X)) if mean_init == "rand" else np.full((Y, X), 1.82)
        mean = rng.uniform(0.0, 1.0, size=(Y, X)) if mean_init == "rand" else np.full((Y, X), 0.5)
        if X == Y:
            acts_tbl = ref_actions.action_dict_to_np_array(
                ref_actions.enumerate_actions(gm, cfg.min_altitude, cfg.max_altitude, cfg.altitude_spacing)
            )
        else:  # the reference's action ids collide on non-square grids (IndexError) -> plain pose list
            acts_tbl = np.array([[res * c + 0.5 * res, res * r + 0.5 * res, h]
                                 for h in orc.altitude_levels(cfg) for r in range(Y) for c in range(X)])
        m_max = X * Y
        prev = np.array([2.0, 2.0, 14.0])
        rec = {k: [] for k in ("action", "prev", "eps", "z", "zshape", "mean", "var", "reward", "reward_adaptive", "action_id")}
        rec["var0"], rec["mean0"], rec["gt"] = var.copy(), mean.copy(), gt
        o_mean, o_var = mean.copy(), var.copy()
        for t in range(T):
            aid = int(rng.randint(len(acts_tbl)))
            a = acts_tbl[aid].copy()
            if t % 5 == 4:  # continuous (off-centre) poses too
                a = np.array([rng.uniform(0, X * res), rng.uniform(0, Y * res), rng.uniform(cfg.min_altitude, cfg.max_altitude)])
                aid = -1
            seed = 9000 + 17 * t
            np.random.seed(seed)
            z = sim.take_measurement(a)
            np.random.seed(seed)
            eps = np.random.standard_normal(z.shape)
            P0 = np.diag(var.ravel())
            gm.mean = mean.copy()
            adaptive_info = {"mean": mean.copy(), "value_threshold": cfg.value_threshold, "interval_factor": 0.25}
            r_plain, _, P1c = simulate_prediction_step(P0, prev, a, mapping, uav, None)
            r_adapt, _, _ = simulate_prediction_step(P0, prev, a, mapping, uav, adaptive_info)
            x1, P1 = mapping.update_grid_map(a, z, cov_only=False, predict_only=True, current_cov_matrix=P0)
            assert np.allclose(np.diag(P1), np.diag(P1c), atol=1e-14)
            mean_n, var_n = np.array(x1), np.diag(P1).reshape(Y, X).copy()
            # ---- oracle check
            cfg_a = copy.copy(cfg)
            cfg_a.interval_factor = 0.25
            ro, om, ov, oz = orc.full_step(cfg, gt, mean, var, prev, a, eps)
            ra, _, _, _ = orc.full_step(cfg_a, gt, mean, var, prev, a, eps, adaptive=True)
            note(f"ep_{name}_z", np.max(np.abs(oz - z)))
            note(f"ep_{name}_mean", np.max(np.abs(om - mean_n)))
            note(f"ep_{name}_var", np.max(np.abs(ov - var_n)))
            note(f"ep_{name}_reward", abs(ro - r_plain) / max(1.0, abs(r_plain)))
            note(f"ep_{name}_reward_adaptive", abs(ra - r_adapt) / max(1.0, abs(r_adapt)))
            zpad = np.zeros(m_max)
            zpad[: z.size] = z.ravel()
            epad = np.zeros(m_max)
            epad[: z.size] = eps.ravel()
            for k, v in (("action", a), ("prev", prev.copy()), ("eps", epad), ("z", zpad), ("zshape", np.array(z.shape)),
                         ("mean", mean_n), ("var", var_n), ("reward", r_plain), ("reward_adaptive", r_adapt), ("action_id", aid if X == Y else -1)):
                rec[k].append(v)
            mean, var, prev = mean_n, var_n, a
        out[f"ep_{name}_cfg"] = cfg_json(p)
        for k, v in rec.items():
            out[f"ep_{name}_{k}"] = np.array(v)
    out["ep_names"] = np.array([s[0] for s in scen])


# ----------------------------------------------------------------------------------------------
def gen_windowed(out):
    """Tier T2: the reference's H builder + static kalman_filter_update on the FoV window of a
    200x200 / 400x400 map with a diagonal covariance; z from the reference's take_measurement."""
    rng = np.random.RandomState(0)
    for name, n in (("w200", 200), ("w400", 400)):
        p = make_params(n, n, 1, 8, 20, 6)
        gm = GridMap(p)
        model = SensorModelFactory(p).create_sensor_model()
        sensor = SensorFactory(p, model, gm).create_sensor()
        np.random.seed(1000 + n)
        sim = SimulationFactory(p, sensor).create_sensor_simulation()
        sensor.set_sensor_simulation(sim)
        cfg = orc.OracleConfig.from_params(p)
        gt = np.array(sim.ground_truth_map)
        var = np.random.RandomState(0).uniform(0.1, 2.0, (n, n))
        mean = np.random.RandomState(1).uniform(0.0, 1.0, (n, n))
        poses = [[n / 2 + 0.5, n / 2 + 0.5, 8], [n / 2 + 0.5, n / 2 + 0.5, 14], [n / 2 + 0.5, n / 2 + 0.5, 20],
                 [3.5, n - 2.5, 20], [0.5, 0.5, 14], [n - 0.5, 5.5, 20], [7.5, n - 0.5, 14], [n - 0.5, n - 0.5, 20]]
        for _ in range(24):
            poses.append([rng.uniform(0, n), rng.uniform(0, n), [8, 14, 20][rng.randint(3)]])
        poses = np.array(poses, dtype=np.float64)
        K = len(poses)
        wmax = 23
        rec = dict(fov=[], tr=[], eps=[], z=[], zshape=[], mean_w=np.zeros((K, wmax, wmax)), var_w=np.zeros((K, wmax, wmax)))
        for k, q in enumerate(poses):
            xl, xr, yu, yd = sensor.project_field_of_view(q)
            rf = sensor.get_resolution_factor(q)
            nx, ny = xr - xl + 1, yd - yu + 1
            wparams = make_params(nx, ny, 1, 8, 20, 6)
            wgm = GridMap(wparams)
            m = int(np.ceil(nx / rf) * np.ceil(ny / rf))
            H = model.measurement_model_matrix(wgm, (0, nx - 1, 0, ny - 1), m, rf)
            R = model.measurement_variance_matrix(q, m, rf)
            np.random.seed(77 + k)
            z = sim.take_measurement(q)
            np.random.seed(77 + k)
            eps = np.random.standard_normal(z.shape)
            vw = var[yu : yd + 1, xl : xr + 1]
            mw = mean[yu : yd + 1, xl : xr + 1]
            x1, P1 = Mapping.kalman_filter_update(np.diag(vw.ravel()), H, R, grid_mean=mw, observation=z, cov_only=False)
            v1 = np.diag(P1).reshape(ny, nx)
            m1 = x1.reshape(ny, nx)
            # oracle check
            _, om, ov, oz = orc.full_step(cfg, gt, mean, var, q, q, eps)
            note(f"{name}_z", np.max(np.abs(oz - z)))
            note(f"{name}_var", np.max(np.abs(ov[yu : yd + 1, xl : xr + 1] - v1)))
            note(f"{name}_mean", np.max(np.abs(om[yu : yd + 1, xl : xr + 1] - m1)))
            chk = ov.copy()
            chk[yu : yd + 1, xl : xr + 1] = var[yu : yd + 1, xl : xr + 1]
            assert np.array_equal(chk, var)  # untouched outside the window
            rec["fov"].append([xl, xr, yu, yd])
            rec["tr"].append(float(np.sum(vw) - np.sum(v1)))
            e = np.zeros(wmax * wmax)
            e[: z.size] = eps.ravel()
            zz = np.zeros(wmax * wmax)
            zz[: z.size] = z.ravel()
            rec["eps"].append(e)
            rec["z"].append(zz)
            rec["zshape"].append(z.shape)
            rec["mean_w"][k, :ny, :nx] = m1
            rec["var_w"][k, :ny, :nx] = v1
        out[f"{name}_cfg"] = cfg_json(p)
        out[f"{name}_gt"] = gt.astype(np.float32)  # engine state is fp32; keep fixtures small
        out[f"{name}_poses"] = poses
        for k, v in rec.items():
            out[f"{name}_{k}"] = np.array(v)
        # NOTE: z/mean were produced from the fp64 GT; tests that feed the fp32 GT use 1e-6 slack.
    # Appendix-B known answers (var = RandomState(0).uniform(0.1, 2.0)) — trace reductions
    out["w200_known_tr"] = np.array([80.186342330, 46.503629155, 80.663532417, 31.474857172, 16.193481081])


# ----------------------------------------------------------------------------------------------
def gen_dense_T0(out):
    """Tier T0: unmodified dense reference with the stock GP prior (example.yaml), plus the same
    calls from a diagonal prior — documents the dense-vs-diagonal gap."""
    p = copy.deepcopy(BASE)
    gm, sensor, sim, mapping = build(p, 0)
    uav = p["experiment"]["uav"]
    prev = np.array([2.0, 2.0, 14.0])
    acts = np.array([[2, 2, 14], [18, 22, 8], [38, 38, 14], [20, 20, 8]], dtype=np.float64)
    dense, diag = [], []
    P_gp = gm.cov_matrix
    P_dg = np.diag(np.diag(P_gp))
    for a in acts:
        r, _, P1 = simulate_prediction_step(P_gp, prev, a, mapping, uav, None)
        dense.append([r, np.trace(P1)])
        r2, _, P2 = simulate_prediction_step(P_dg, prev, a, mapping, uav, None)
        diag.append([r2, np.trace(P2)])
    out["t0_actions"] = acts
    out["t0_prev"] = prev
    out["t0_dense_reward_trace"] = np.array(dense)
    out["t0_diag_reward_trace"] = np.array(diag)
    out["t0_prior_diag"] = np.diag(P_gp).copy()
    out["t0_cfg"] = cfg_json(p)


# ----------------------------------------------------------------------------------------------
def gen_actions_metrics(out):
    p = copy.deepcopy(BASE)
    gm = GridMap(p)
    cfg = orc.OracleConfig.from_params(p)
    tbl = ref_actions.action_dict_to_np_array(ref_actions.enumerate_actions(gm, 8, 14, 6))
    assert np.array_equal(tbl, orc.enumerate_actions(cfg))
    out["act_ex10_table"] = tbl
    p6 = make_params(6, 6, 2.5, 5, 15, 5)
    tbl6 = ref_actions.action_dict_to_np_array(ref_actions.enumerate_actions(GridMap(p6), 5, 15, 5))
    assert np.array_equal(tbl6, orc.enumerate_actions(orc.OracleConfig.from_params(p6)))
    out["act_g6_table"] = tbl6
    out["act_g6_cfg"] = cfg_json(p6)
    rng = np.random.RandomState(3)
    a = rng.uniform(0, 40, (64, 3))
    b = rng.uniform(0, 40, (64, 3))
    b[:8] = a[:8]  # zero distance
    b[8:16] = a[8:16] + rng.uniform(-0.5, 0.5, (8, 3))  # shorter than the acceleration distance
    uav = p["experiment"]["uav"]
    ft = np.array([ref_actions.action_costs(x, y, uav) for x, y in zip(a, b)])
    ed = np.array([ref_actions.action_costs(x, y, None) for x, y in zip(a, b)])
    for x, y, f, e in zip(a, b, ft, ed):
        note("flight_time", abs(orc.action_costs(x, y, uav) - f))
        note("distance", abs(orc.action_costs(x, y, None) - e))
    out["cost_a"], out["cost_b"], out["cost_flight_time"], out["cost_distance"] = a, b, ft, ed
    # evaluation metrics
    mets = []
    gts, means, vars_, masks = [], [], [], []
    for k in range(6):
        Y, X = (12, 12) if k < 4 else (9, 14)
        gt = rng.uniform(0, 1, (Y, X))
        mean = np.clip(gt + rng.normal(0, 0.2, (Y, X)), 0, 1) if k % 2 else np.full((Y, X), 0.5) + rng.normal(0, 0.01, (Y, X))
        var = rng.uniform(0.01, 2.0, (Y, X))
        msk = (mean + 0.3 * var >= 0.6)
        P = np.diag(var.ravel())
        mflat = msk.ravel()
        with np.errstate(all="ignore"):
            row = [
                ref_metrics.root_mean_squared_error(gt, mean),
                ref_metrics.weighted_root_mean_squared_error(gt, mean),
                ref_metrics.mean_log_loss(gt, mean, P),
                ref_metrics.weighted_mean_log_loss(gt, mean, P),
                ref_metrics.map_uncertainty(P),
                ref_metrics.map_uncertainty_difference(P, mflat),
                ref_metrics.root_mean_squared_error(gt, mean, mflat),
                ref_metrics.map_uncertainty(P, mflat),
            ]
            o = orc.evaluation_metrics(gt, mean, var, msk)
        for i, (x, y) in enumerate(zip(row, o)):
            if np.isnan(x):
                assert np.isnan(y), (k, i)
            else:
                note(f"metric_{orc.METRIC_NAMES[i]}", abs(x - y) / max(1.0, abs(x)))
        mets.append(row)
        pad = lambda arr: np.pad(arr, ((0, 12 - arr.shape[0]), (0, 14 - arr.shape[1])))  # noqa: E731
        gts.append(pad(gt)), means.append(pad(mean)), vars_.append(pad(var)), masks.append(pad(msk))
    out["met_values"] = np.array(mets)
    out["met_gt"], out["met_mean"], out["met_var"], out["met_mask"] = map(np.array, (gts, means, vars_, masks))
    out["met_shapes"] = np.array([(12, 12)] * 4 + [(9, 14)] * 2)


def main():
    a, b, c = {}, {}, {}
    gen_footprints(a)
    gen_actions_metrics(a)
    gen_dense_T0(a)
    gen_episodes(b)
    gen_windowed(c)
    np.savez_compressed(os.path.join(HERE, "golden_sensor_actions_metrics.npz"), **a)
    np.savez_compressed(os.path.join(HERE, "golden_episodes_T1.npz"), **b)
    np.savez_compressed(os.path.join(HERE, "golden_windowed_T2.npz"), **c)
    meta = {
        "reference": "dmar-bonn/ipp-rl @ 25dfb33 (/root/reference), imported in place with stub plotting modules",
        "numpy": np.__version__,
        "worst_abs_or_rel_deviation_oracle_vs_reference": WORST,
    }
    import cv2

    meta["cv2"] = cv2.__version__
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for k in sorted(WORST):
        print(f"{k:40s} {WORST[k]:.3e}")


if __name__ == "__main__":
    main()
