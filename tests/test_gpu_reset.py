"""GPU tests of the reset path: on-device Gaussian-random-field ground truth (ipp_generate_ground_truth) against the REAL
reference's fields (tests/golden/golden_grf.npz, same white-noise draw), and the device-noise mode's properties."""
import numpy as np
import pytest

from tests._util import engine_cfg, golden, make_params

pytestmark = pytest.mark.gpu


def _engine(params, batch, **kw):
    from ipp_rl_b200 import BatchedEngine

    return BatchedEngine(engine_cfg(params, batch, **kw))


@pytest.mark.parametrize("layout", [0, 1, 2, 3, 4])
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


@pytest.mark.parametrize("layout", [0, 1, 2, 3, 4])
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


@pytest.mark.parametrize("layout", [1, 3, 4])
def test_device_hotspot_and_split_fields(layout):
    """ipp_generate_field: the reference's construction (simulations/simulations.py:57-125) per env on the device — two-valued
    maps with the reference's value ranges and geometry, keyed by the global env id (sharding invariance), and the same
    statistics as the host twins that are pinned bit for bit against the reference (tests/test_missions_host.py)."""
    X, Y, r, B = 40, 30, 5, 512
    params = make_params(X, Y, 1.0, 8, 14, 6)
    with _engine(params, B, layout=layout, seed=1) as eng:
        eng.generate_field("hotspot_random_field", r, seed=11)
        hot = eng.get_ground_truth()
        eng.generate_field("split_random_field", r, seed=12)
        split = eng.get_ground_truth()
    with _engine(params, 64, layout=1, seed=1, env_id_offset=100) as eng:
        eng.generate_field("hotspot_random_field", r, seed=11)
        assert np.array_equal(eng.get_ground_truth(), hot[100:164])  # keyed by the GLOBAL env id
        eng.generate_field("split_random_field", r, seed=12, first_env=10, n_env=20)
        assert np.array_equal(eng.get_ground_truth(10, 20), split[110:130])
    areas = []
    for b in range(B):
        vals = np.unique(hot[b])
        assert len(vals) == 2 and 0.0 <= vals[0] <= 0.3 and 0.7 <= vals[1] <= 1.0
        m = hot[b] == vals[1]
        rows, cols = np.flatnonzero(m.any(1)), np.flatnonzero(m.any(0))
        # two axis-aligned squares of side <= 2r whose centres are more than r apart in BOTH axes: their row ranges may touch, never nest
        assert m.sum() <= 2 * (2 * r) ** 2 and m.sum() >= r * r
        areas.append(m.sum())
        lab_rows = np.split(rows, np.flatnonzero(np.diff(rows) > 1) + 1)
        assert len(lab_rows) <= 2 and len(np.split(cols, np.flatnonzero(np.diff(cols) > 1) + 1)) <= 2
        vs = np.unique(split[b])
        assert len(vs) == 2 and 0.0 <= vs[0] <= 0.35 and 0.65 <= vs[1] <= 1.0
        first = split[b][0, 0]
        by_rows = np.all(split[b] == split[b][:, :1])  # constant along x -> split along y
        if by_rows:
            cut = int(np.argmax(split[b][:, 0] != first))
            assert int(np.ceil(Y * 0.33)) <= cut <= int(np.ceil(Y * 0.66)) and np.all(split[b][:cut] == first) and np.all(split[b][cut:] != first)
        else:
            assert np.all(split[b] == split[b][:1, :])
            cut = int(np.argmax(split[b][0] != first))
            assert int(np.floor(X * 0.33)) <= cut <= int(np.ceil(X * 0.66)) and np.all(split[b][:, :cut] == first)
    # statistics against the (reference-pinned) host twins under NumPy's RNG
    from ipp_rl_b200.mapping.grid_maps import GridMap
    from ipp_rl_b200.sensors.models.sensor_model_factories import SensorModelFactory
    from ipp_rl_b200.sensors.sensor_factories import SensorFactory
    from ipp_rl_b200.simulations.simulations import HotspotRandomField, SplitRandomField

    gm = GridMap(params)
    sensor = SensorFactory(params, SensorModelFactory(params).create_sensor_model(), gm).create_sensor()
    np.random.seed(0)
    host_hot = np.stack([HotspotRandomField(sensor, r).ground_truth_map for _ in range(B)])
    host_split = np.stack([SplitRandomField(sensor, r).ground_truth_map for _ in range(B)])
    assert abs(hot.mean() - host_hot.mean()) < 0.03 and abs(split.mean() - host_split.mean()) < 0.05
    assert abs(np.mean(areas) - np.mean([(h == h.max()).sum() for h in host_hot])) < 12
    by_rows_dev = np.mean([np.all(s == s[:, :1]) for s in split])
    assert 0.4 < by_rows_dev < 0.6


@pytest.mark.parametrize("layout", [0, 3, 4])
def test_device_shuffled_priors(layout):
    """ipp_reset_shuffled: Mapping.init_priors(shuffle_prior_cov=True) per env (mapping/mappings.py:219-240)."""
    X = Y = 20
    B = 1024
    params = make_params(X, Y, 1.0, 8, 14, 6)
    with _engine(params, B, layout=layout, seed=1) as eng:
        eng.reset_shuffled(0.5, fit_gaussian_process=True, scale=1.82, seed=5)
        m, v = eng.get_state()
        assert np.all(m == 0.5)
        lv = v.reshape(B, -1)
        assert np.all(lv == lv[:, :1])  # GP mode: the Matern prior's diagonal is one signal variance per env
        assert lv[:, 0].min() >= 0.8 * 1.82 - 1e-6 and lv[:, 0].max() <= 1.2 * 1.82 + 1e-6
        assert abs(lv[:, 0].mean() - 1.82) < 0.02 and abs(lv[:, 0].std() - 0.4 * 1.82 / np.sqrt(12)) < 0.02
        assert np.allclose(eng.get_prev_pose(), [2.0, 2.0, 14.0])
        eng.reset_shuffled(0.5, fit_gaussian_process=False, scale=0.5, seed=6)
        _, v2 = eng.get_state()
    with _engine(params, 32, layout=1, seed=1, env_id_offset=64) as eng:
        eng.reset_shuffled(0.5, fit_gaussian_process=True, scale=1.82, seed=5)
        assert np.array_equal(eng.get_state()[1], v[64:96])
    # non-GP: diag(A A^T) / ||A||_F with A ~ N(mu, mu), mu ~ U(0.1, 0.5): level sqrt(2) mu, a small per-cell spread
    lv2 = v2.reshape(B, -1)
    level = lv2.mean(1)
    assert level.min() >= np.sqrt(2) * 0.1 * 0.97 and level.max() <= np.sqrt(2) * 0.5 * 1.03
    assert abs(level.mean() - np.sqrt(2) * 0.3) < 0.02
    rel = lv2.std(1) / level
    assert np.all(rel > 0) and abs(rel.mean() - np.sqrt(6.0 / (X * Y)) / 2) < 0.01
    # the same statistic from the reference's construction (host, NumPy)
    rng = np.random.RandomState(0)
    a = rng.normal(0.3, 0.3, (X * Y, X * Y))
    d = np.einsum("ij,ij->i", a, a) / np.linalg.norm(a)
    assert abs(d.mean() - np.sqrt(2) * 0.3) < 0.01 and abs(d.std() / d.mean() - np.sqrt(6.0 / (X * Y)) / 2) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1, 2, 3, 4])
def test_observation_planes_vector_path_equals_generic_path(layout, monkeypatch):
    """The 16-byte path of ipp_observe (x_dim % 4 == 0) writes the bits of the one-cell-per-thread kernel: all six planes, with
    and without the adaptive mask, poses on and off the lattice."""
    params = make_params(48, 40, 1.0, 8, 20, 6, kappa=0.3, thr=0.5)
    B = 6
    rng = np.random.RandomState(3)
    mean0 = rng.uniform(0, 1, (B, 40, 48)).astype(np.float32)
    var0 = rng.uniform(0.05, 2.0, (B, 40, 48)).astype(np.float32)
    poses = np.stack([rng.uniform(0, 48, B), rng.uniform(0, 40, B), rng.choice([8.0, 14.0, 20.0], B)], axis=1)
    budget = rng.uniform(0.1, 1.0, B).astype(np.float32)
    with _engine(params, B, layout=layout) as eng:
        eng.reset(0.5, 1.0)
        eng.set_state(mean0, var0)
        for adaptive in (False, True):
            fast = eng.observe(0, B, poses=poses, budget_ratio=budget, adaptive=adaptive)
            monkeypatch.setenv("IPP_OBS_GENERIC", "1")
            slow = eng.observe(0, B, poses=poses, budget_ratio=budget, adaptive=adaptive)
            monkeypatch.delenv("IPP_OBS_GENERIC")
            assert fast.shape == (B, 6, 40, 48) and np.array_equal(fast, slow), adaptive
            five = eng.observe(0, B, poses=poses, budget_ratio=budget, adaptive=adaptive, action_costs=False)
            assert np.array_equal(five, fast[:, :5])
