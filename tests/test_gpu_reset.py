"""GPU tests of the reset path: on-device Gaussian-random-field ground truth (ipp_generate_ground_truth) against the REAL
reference's fields (tests/golden/golden_grf.npz, same white-noise draw), and the device-noise mode's properties."""
import numpy as np
import pytest

from tests._util import engine_cfg, golden, make_params

pytestmark = pytest.mark.gpu


def _engine(params, batch, **kw):
    from ipp_rl_b200 import BatchedEngine

    return BatchedEngine(engine_cfg(params, batch, **kw))


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_device_grf_matches_reference_fields(layout):
    g = golden("golden_grf.npz")
    for name in g["names"]:
        X, Y = (int(v) for v in g[f"{name}_dims"])
        params = make_params(X, Y, 1.0, 8, 14, 6)
        white = g[f"{name}_white"]
        with _engine(params, 3, layout=layout) as eng:
            eng.generate_ground_truth(float(g[f"{name}_radius"]), white_noise=np.stack([white, -white, white]))
            gt = eng.get_ground_truth()
        ref = g[f"{name}_field"]
        assert gt.min() == 0.0 and gt.max() == 1.0
        assert np.max(np.abs(gt[0] - ref)) <= 2e-4, (name, np.max(np.abs(gt[0] - ref)))
        assert np.array_equal(gt[0], gt[2])
        assert np.max(np.abs(gt[1] - (1.0 - ref))) <= 2e-4  # the generator is linear before the min-max normalisation


def test_device_grf_philox_mode_is_shard_invariant_and_smooth():
    X = Y = 200
    params = make_params(X, Y, 1.0, 8, 20, 6)
    with _engine(params, 64, layout=2, seed=1) as eng:
        eng.generate_ground_truth(5.0, seed=99)
        full = eng.get_ground_truth()
    with _engine(params, 16, layout=1, seed=1, env_id_offset=32) as eng:
        eng.generate_ground_truth(5.0, seed=99)
        part = eng.get_ground_truth()
    assert np.max(np.abs(part - full[32:48])) <= 2e-6  # same (seed, global env id) -> same field, up to FFT batching
    assert np.all(full.min(axis=(1, 2)) == 0.0) and np.all(full.max(axis=(1, 2)) == 1.0)
    assert len({full[k].tobytes() for k in range(64)}) == 64  # every env its own field
    f = full - full.mean(axis=(1, 2), keepdims=True)
    corr = (f[:, :, 1:] * f[:, :, :-1]).sum(axis=(1, 2)) / (f * f).sum(axis=(1, 2))
    assert np.all(corr > 0.9)  # k^-5 spectrum: strongly correlated neighbours (white noise would give ~0)


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_observation_planes_match_reference_feature_planes(layout):
    """ipp_observe vs the REAL reference generate_input_feature_planes on a diagonal state (golden_features.npz)."""
    from tests._util import params_from_json

    g = golden("golden_features.npz")
    for name in g["names"]:
        params = params_from_json(g[f"{name}_cfg"])
        X, Y = params["environment"]["x_dim"], params["environment"]["y_dim"]
        for adaptive in (False, True):
            tag = f"{name}_{'adaptive' if adaptive else 'plain'}"
            with _engine(params, 2, layout=layout, value_threshold=0.5, interval_factor=0.3) as eng:
                eng.reset(0.5, 1.0)
                eng.set_state(np.broadcast_to(g[f"{name}_mean"], (2, Y, X)), np.broadcast_to(g[f"{name}_var"], (2, Y, X)))
                obs = eng.observe(poses=g[f"{name}_pos"], budget_ratio=float(g[f"{name}_ratio"]), adaptive=adaptive, action_costs=True)
            assert obs.shape == (2, 6, Y, X)
            assert np.array_equal(obs[0], obs[1])
            ref_state = g[f"{tag}_state_diag"]
            # a mask decision within fp32 rounding of the threshold may differ: skip cells that close to it
            near = np.abs(g[f"{name}_mean"] + 0.3 * g[f"{name}_var"] - 0.5) < 1e-5 if adaptive else np.zeros((Y, X), bool)
            assert np.max(np.abs(obs[0, 0] - ref_state)[~near]) <= 1e-6
            for k in range(4):
                assert np.all(np.abs(obs[0, 1 + k] - g[f"{tag}_consts"][k]) <= 1e-6)
            cost = g[f"{tag}_cost_by_action"]  # indexed by action id X*col + row (planning/common/actions.py:94-96)
            cols, rows = np.meshgrid(np.arange(X), np.arange(Y))
            assert np.max(np.abs(obs[0, 5] - cost[X * cols + rows])) <= 1e-6
