"""Shared helpers for the test-suite (golden loading, oracle/engine config construction)."""
import json
import os

import numpy as np

from oracle import ipp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return _cache[name]


def params_from_json(s) -> dict:
    p = json.loads(str(s))
    return p


def oracle_cfg(params: dict, **over) -> orc.OracleConfig:
    c = orc.OracleConfig.from_params(params)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def engine_cfg(params: dict, batch: int, **over):
    from ipp_rl_b200 import EngineConfig

    return EngineConfig.from_params(params, batch=batch, **over)


def make_params(x_dim, y_dim, res, alt_min, alt_max, alt_step, angle=(60.0, 60.0), thr=0.4, kappa=0.0, uav=(2.0, 2.0)):
    p = {
        "environment": {"x_dim": x_dim, "y_dim": y_dim, "resolution": res},
        "sensor": {
            "type": "rgb_camera",
            "field_of_view": {"angle_x": angle[0], "angle_y": angle[1]},
            "encoding": "rgb8",
            "model": {"type": "altitude_dependent", "coeff_a": 0.05, "coeff_b": 0.2},
            "simulation": {"type": "gaussian_random_field", "cluster_radius": 5},
        },
        "mapping": {"fit_gaussian_process": False, "prior_cov_mean": 0.5, "prior_cov_std": 0.25, "signal_variance": 1.82,
                    "length_scale": 3.67, "noise_variance": 1.42, "nu": 1.5},
        "experiment": {
            "constraints": {"min_altitude": alt_min, "max_altitude": alt_max, "altitude_spacing": alt_step, "budget": 200,
                            "dist_to_boundaries": 3},
            "scenario": {"adaptive": True, "value_threshold": thr, "interval_factor": kappa},
        },
    }
    if uav is not None:
        p["experiment"]["uav"] = {"max_v": uav[0], "max_a": uav[1], "sampling_time": 2}
    return p


def smooth_field(rng, shape):
    """Cheap smooth random field in [0,1] (test input only)."""
    Y, X = shape
    yy, xx = np.mgrid[0:Y, 0:X]
    f = np.zeros(shape)
    for k in range(5):
        kx, ky = rng.uniform(-0.3, 0.3, 2)
        f += rng.uniform(0.3, 1.0) * np.sin(kx * xx + ky * yy + rng.uniform(0, 6.28))
    f = (f - f.min()) / (f.max() - f.min())
    return f


def stub_policy_value(prev, budget_ratio, num_actions):
    """The deterministic stand-in for the policy/value network that tests/golden/make_golden_mcts.py used when it ran the
    reference MCTS (same function, duplicated here because that script imports /root/reference)."""
    k = int(round(prev[0] + 7 * prev[1] + 13 * prev[2])) + int(round(100 * budget_ratio))
    a = np.arange(num_actions, dtype=np.float64)
    s = np.sin(a * 12.9898 + k * 78.233) * 43758.5453
    policy = 0.2 + (s - np.floor(s))
    return policy / policy.sum(), 0.05 * (k % 7)
