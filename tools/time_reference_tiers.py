#!/usr/bin/env python
"""Time the REFERENCE'S OWN CODE on the hot path (tiers T0 and T2) in the build container.

The reference is pure Python and does not exist on the GPU box (it cannot travel; only golden vectors do), so these tiers are
timed here, where /root/reference is, and the result is committed as profiles/r02_reference_cpu_tiers.json; bench.py attaches
it to ``cpu_baseline`` next to the port timed on the GPU box's cores.

  T0  simulate_prediction_step (planning/common/optimization.py:14-30) and take_measurement + update_grid_map
      (simulations/simulations.py:26-34, mapping/mappings.py:114-215) with the dense GP prior at 10x10 and 50x50
  T2  the reference's measurement_model_matrix + static kalman_filter_update on the FoV window of a 200x200 map with a
      diagonal covariance — the reference's arithmetic at the target size (a dense 200x200 env would need 12.8 GB)

    PYTHONDONTWRITEBYTECODE=1 python tools/time_reference_tiers.py
"""
import copy
import json
import os
import sys
import time
import types

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
for _n in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "cma", "telegram"]:
    sys.modules[_n] = types.ModuleType(_n)
sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

import numpy as np  # noqa: E402
import yaml  # noqa: E402

from mapping.grid_maps import GridMap  # noqa: E402
from mapping.mappings import Mapping  # noqa: E402
from planning.common.optimization import simulate_prediction_step  # noqa: E402
from sensors.models.sensor_model_factories import SensorModelFactory  # noqa: E402
from sensors.sensor_factories import SensorFactory  # noqa: E402
from simulations.simulation_factories import SimulationFactory  # noqa: E402

BASE = yaml.safe_load(open(os.path.join(REF, "config/example.yaml")))


def build(params, seed, with_mapping=True):
    np.random.seed(seed)
    gm = GridMap(params)
    model = SensorModelFactory(params).create_sensor_model()
    sensor = SensorFactory(params, model, gm).create_sensor()
    sim = SimulationFactory(params, sensor).create_sensor_simulation()
    sensor.set_sensor_simulation(sim)
    return gm, model, sensor, sim, (Mapping(gm, sensor) if with_mapping else None)


def t0(n, res, reps):
    p = copy.deepcopy(BASE)
    p["environment"].update(x_dim=n, y_dim=n, resolution=res)
    gm, model, sensor, sim, mapping = build(p, 1)
    uav = p["experiment"]["uav"]
    rng = np.random.RandomState(0)
    alts = [8.0, 14.0]
    acts = [np.array([res * rng.randint(n) + res / 2, res * rng.randint(n) + res / 2, alts[rng.randint(2)]]) for _ in range(reps)]
    prev = np.array([2.0, 2.0, 14.0])
    P = mapping.grid_map.cov_matrix
    t = time.perf_counter()
    for a in acts:
        simulate_prediction_step(P, prev, a, mapping, uav, None)
    predict_ms = 1e3 * (time.perf_counter() - t) / reps
    t = time.perf_counter()
    for a in acts:
        z = sensor.take_measurement(a, verbose=False)
        mapping.update_grid_map(a, z)
    full_ms = 1e3 * (time.perf_counter() - t) / reps
    return {"grid": [n, n], "reps": reps, "simulate_prediction_step_ms": predict_ms, "take_measurement_plus_update_grid_map_ms": full_ms,
            "steps_per_sec_one_process": 1e3 / full_ms}


def t2(n, reps):
    p = copy.deepcopy(BASE)
    p["environment"].update(x_dim=n, y_dim=n, resolution=1)
    p["experiment"]["constraints"].update(min_altitude=8, max_altitude=20, altitude_spacing=6)
    gm, model, sensor, sim, _ = build(p, 2, with_mapping=False)
    rng = np.random.RandomState(0)
    var = rng.uniform(0.1, 2.0, (n, n))
    mean = rng.uniform(0.0, 1.0, (n, n))
    poses = [np.array([rng.randint(n) + 0.5, rng.randint(n) + 0.5, [8.0, 14.0, 20.0][rng.randint(3)]]) for _ in range(reps)]
    t = time.perf_counter()
    for q in poses:
        xl, xr, yu, yd = sensor.project_field_of_view(q)
        rf = sensor.get_resolution_factor(q)
        nx, ny = xr - xl + 1, yd - yu + 1
        wp = copy.deepcopy(p)
        wp["environment"].update(x_dim=nx, y_dim=ny)
        wgm = GridMap(wp)
        m = int(np.ceil(nx / rf) * np.ceil(ny / rf))
        H = model.measurement_model_matrix(wgm, (0, nx - 1, 0, ny - 1), m, rf)
        R = model.measurement_variance_matrix(q, m, rf)
        z = sim.take_measurement(q)
        vw, mw = var[yu : yd + 1, xl : xr + 1], mean[yu : yd + 1, xl : xr + 1]
        x1, P1 = Mapping.kalman_filter_update(np.diag(vw.ravel()), H, R, grid_mean=mw, observation=z, cov_only=False)
        var[yu : yd + 1, xl : xr + 1] = np.diag(P1).reshape(ny, nx)
        mean[yu : yd + 1, xl : xr + 1] = x1.reshape(ny, nx)
    ms = 1e3 * (time.perf_counter() - t) / reps
    return {"grid": [n, n], "reps": reps, "windowed_full_step_ms": ms, "steps_per_sec_one_process": 1e3 / ms}


if __name__ == "__main__":
    out = {"where": "build container (no GPU): the reference is pure Python and cannot travel to the GPU box", "cpu_count": os.cpu_count(),
           "numpy": np.__version__, "kind": "reference",
           "T0_dense_reference": [t0(10, 4, 300), t0(50, 4, 20)],
           "T2_reference_on_window": [t2(200, 300)],
           "note": "one Python process, default BLAS threads; the dense reference cannot run at 200x200 (12.8 GB fp64 covariance per env)"}
    path = os.path.join(ROOT, "profiles", "r02_reference_cpu_tiers.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
