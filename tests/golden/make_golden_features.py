#!/usr/bin/env python
"""Golden vectors for the observation planes, produced by RUNNING THE REAL REFERENCE
``generate_input_feature_planes`` (planning/common/features.py:83-151) on a DIAGONAL state, history length 1:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_features.py

Stored per case: the diagonal of the N x N state plane, the four constant planes' values, and the cost plane's first
column (the plane is constant along rows) — what ipp_observe emits as (y_dim, x_dim) planes."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (reference import path + stub modules)

import numpy as np  # noqa: E402
from planning.common.features import EpisodeHistory, generate_input_feature_planes  # noqa: E402  (reference)


def main():
    out = {}
    cases = [("a", 10, 10, 4.0, 8, 14, 6, [18.0, 22.0, 14.0], 0.6), ("b", 12, 12, 2.0, 6, 18, 6, [3.0, 21.0, 12.0], 0.25)]
    for name, X, Y, res, a0, a1, da, pos, ratio in cases:
        params = mg.make_params(X, Y, res, a0, a1, da)
        gm, sensor, sim, mapping = mg.build(params, 3)
        rng = np.random.RandomState(17)
        var = rng.uniform(0.05, 2.0, (Y, X))
        mean = rng.uniform(0.0, 1.0, (Y, X))
        gm.mean = mean.copy()
        uav = params["experiment"]["uav"]
        for adaptive in (False, True):
            hist = EpisodeHistory(1)
            hist.push(np.diag(var.flatten()).copy(), np.array(pos, float), ratio)
            info = {"mean": mean.copy(), "value_threshold": 0.5, "interval_factor": 0.3} if adaptive else None
            planes = generate_input_feature_planes(mapping, hist, a0, a1, adaptive_info=info, uav_specifications=uav,
                                                   use_action_costs_input=True)
            assert planes.shape == (6, X * Y, X * Y)
            tag = f"{name}_{'adaptive' if adaptive else 'plain'}"
            out[f"{tag}_state_diag"] = np.diag(planes[0]).reshape(Y, X).copy()
            out[f"{tag}_consts"] = np.array([planes[k][0, 0] for k in (1, 2, 3, 4)])
            assert all(np.all(planes[k] == planes[k][0, 0]) for k in (1, 2, 3, 4))
            assert np.all(planes[5] == planes[5][:, :1])
            out[f"{tag}_cost_by_action"] = planes[5][:, 0].copy()
        out[f"{name}_cfg"] = mg.cfg_json(params)
        out[f"{name}_var"], out[f"{name}_mean"] = var, mean
        out[f"{name}_pos"], out[f"{name}_ratio"] = np.array(pos), np.array(ratio)
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "golden_features.npz"), **out)
    print("written", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
