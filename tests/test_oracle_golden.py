"""CPU: pin the oracle (NumPy restatement + its C port) against the golden vectors that
tests/golden/make_golden.py produced by running the REAL reference.  No GPU, no /root/reference."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import ipp_oracle as orc
from tests._util import golden, oracle_cfg, params_from_json

G1 = "golden_sensor_actions_metrics.npz"
G2 = "golden_episodes_T1.npz"
G3 = "golden_windowed_T2.npz"


def test_footprint_rf_sigma2_match_reference():
    g = golden(G1)
    for si in range(int(g["fp_count"])):
        cfg = oracle_cfg(params_from_json(g[f"fp{si}_cfg"]))
        poses, fov, rf = g[f"fp{si}_poses"], g[f"fp{si}_fov"], g[f"fp{si}_rf"]
        for q, f, r, s2, R in zip(poses, fov, rf, g[f"fp{si}_sigma2"], g[f"fp{si}_R"]):
            assert orc.project_field_of_view(cfg, q) == tuple(int(t) for t in f)
            assert orc.resolution_factor(cfg, q) == r
            assert orc.noise_variance(cfg, q) == pytest.approx(s2, abs=1e-16)
            assert orc.measurement_variance(cfg, q, int(r)) == pytest.approx(R, abs=1e-16)


def test_measurement_matrix_structure_matches_reference():
    g = golden(G1)
    for si in range(int(g["fp_count"])):
        key = f"fp{si}_Hstats"
        if key not in g:
            continue
        cfg = oracle_cfg(params_from_json(g[f"fp{si}_cfg"]))
        sub = list(zip(g[f"fp{si}_poses"], g[f"fp{si}_fov"], g[f"fp{si}_rf"]))[::7]
        for (q, f, r), stats in zip(sub, g[key]):
            H = orc.measurement_model_matrix(cfg, tuple(int(t) for t in f), int(r))
            assert [H.shape[0], int(np.count_nonzero(H)), float(H.sum())] == pytest.approx(list(stats))


def test_action_table_costs_metrics_match_reference():
    g = golden(G1)
    assert np.array_equal(orc.enumerate_actions(oracle_cfg(params_from_json(g["t0_cfg"]))), g["act_ex10_table"])
    assert np.array_equal(orc.enumerate_actions(oracle_cfg(params_from_json(g["act_g6_cfg"]))), g["act_g6_table"])
    uav = {"max_v": 2, "max_a": 2}
    for a, b, ft, ed in zip(g["cost_a"], g["cost_b"], g["cost_flight_time"], g["cost_distance"]):
        assert orc.action_costs(a, b, uav) == pytest.approx(ft, abs=1e-14)
        assert orc.action_costs(a, b, None) == pytest.approx(ed, abs=1e-14)
    for k, (Y, X) in enumerate(g["met_shapes"]):
        gt, mean, var, msk = (g[n][k, :Y, :X] for n in ("met_gt", "met_mean", "met_var", "met_mask"))
        with np.errstate(all="ignore"):
            o = orc.evaluation_metrics(gt, mean, var, msk.astype(bool))
        for x, y in zip(g["met_values"][k], o):
            assert (np.isnan(x) and np.isnan(y)) or y == pytest.approx(x, rel=1e-12)


def test_dense_reference_known_answers_and_diagonal_gap():
    """the dense GP-prior rewards are reproduced by the reference itself (stored), the
    same calls from diag(P) equal the per-cell oracle — the documented dense-vs-diagonal gap."""
    g = golden(G1)
    known = [10.024260143748165, 1.4140125945387325, 0.3410710697668489, 1.421163889436027]
    assert g["t0_dense_reward_trace"][:, 0] == pytest.approx(known, rel=1e-12)
    cfg = oracle_cfg(params_from_json(g["t0_cfg"]))
    var = g["t0_prior_diag"].reshape(cfg.y_dim, cfg.x_dim)
    for a, (r_ref, tr_ref) in zip(g["t0_actions"], g["t0_diag_reward_trace"]):
        r, v1 = orc.simulate_prediction_step(cfg, var, g["t0_prev"], a)
        assert r == pytest.approx(r_ref, rel=1e-12) and v1.sum() == pytest.approx(tr_ref, rel=1e-12)
    assert np.all(g["t0_diag_reward_trace"][:, 0] < g["t0_dense_reward_trace"][:, 0])  # diag prior carries less information


@pytest.mark.parametrize("name", ["ex10", "g24", "ns30x20", "g16"])
def test_episode_T1_matches_reference(name):
    g = golden(G2)
    params = params_from_json(g[f"ep_{name}_cfg"])
    cfg = oracle_cfg(params)
    cfg_a = oracle_cfg(params, interval_factor=0.25)
    gt = g[f"ep_{name}_gt"]
    mean, var = g[f"ep_{name}_mean0"], g[f"ep_{name}_var0"]
    for t in range(len(g[f"ep_{name}_action"])):
        a, prev = g[f"ep_{name}_action"][t], g[f"ep_{name}_prev"][t]
        m = int(np.prod(g[f"ep_{name}_zshape"][t]))
        eps = g[f"ep_{name}_eps"][t][:m]
        r, mn, vn, z = orc.full_step(cfg, gt, mean, var, prev, a, eps)
        ra, _, _, _ = orc.full_step(cfg_a, gt, mean, var, prev, a, eps, adaptive=True)
        assert z.ravel() == pytest.approx(g[f"ep_{name}_z"][t][:m], abs=1e-7)  # cv2 uses float32 weights
        assert np.max(np.abs(mn - g[f"ep_{name}_mean"][t])) < 1e-7
        assert np.max(np.abs(vn - g[f"ep_{name}_var"][t])) < 1e-13
        assert r == pytest.approx(g[f"ep_{name}_reward"][t], rel=1e-12)
        assert ra == pytest.approx(g[f"ep_{name}_reward_adaptive"][t], rel=1e-12)
        if g[f"ep_{name}_action_id"][t] >= 0:
            assert np.array_equal(orc.enumerate_actions(cfg)[g[f"ep_{name}_action_id"][t]], a)
        mean, var = g[f"ep_{name}_mean"][t], g[f"ep_{name}_var"][t]


@pytest.mark.parametrize("name", ["w200", "w400"])
def test_windowed_T2_matches_reference(name):
    g = golden(G3)
    cfg = oracle_cfg(params_from_json(g[f"{name}_cfg"]))
    n = cfg.x_dim
    gt = g[f"{name}_gt"].astype(np.float64)  # stored as fp32: z / mean agree to fp32 rounding of the GT
    var = np.random.RandomState(0).uniform(0.1, 2.0, (n, n))
    mean = np.random.RandomState(1).uniform(0.0, 1.0, (n, n))
    for k, q in enumerate(g[f"{name}_poses"]):
        xl, xr, yu, yd = g[f"{name}_fov"][k]
        assert orc.project_field_of_view(cfg, q) == (xl, xr, yu, yd)
        m = int(np.prod(g[f"{name}_zshape"][k]))
        _, mn, vn, z = orc.full_step(cfg, gt, mean, var, q, q, g[f"{name}_eps"][k][:m])
        ny, nx = yd - yu + 1, xr - xl + 1
        assert z.ravel() == pytest.approx(g[f"{name}_z"][k][:m], abs=2e-7)
        assert np.max(np.abs(vn[yu : yd + 1, xl : xr + 1] - g[f"{name}_var_w"][k, :ny, :nx])) < 1e-13
        assert np.max(np.abs(mn[yu : yd + 1, xl : xr + 1] - g[f"{name}_mean_w"][k, :ny, :nx])) < 5e-7
        assert (var - vn).sum() == pytest.approx(g[f"{name}_tr"][k], rel=1e-11)
    if name == "w200":
        assert g["w200_tr"][:5] == pytest.approx(g["w200_known_tr"], rel=1e-9)


def test_inter_area_weights_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for ny, nx in [(17, 17), (23, 23), (9, 9), (5, 5), (14, 23), (23, 12), (15, 14), (12, 17), (3, 5), (5, 3), (4, 3), (2, 2), (1, 1), (16, 16)]:
        img = rng.uniform(0, 1, (ny, nx))
        rows, cols = int(np.ceil(nx / 2)), int(np.ceil(ny / 2))  # dsize swap
        if rows > ny or cols > nx:
            with pytest.raises(NotImplementedError):
                orc.downsample_measurement(img, 2)
            continue
        ref = cv2.resize(img, dsize=(cols, rows), interpolation=cv2.INTER_AREA)
        out = orc.downsample_measurement(img, 2)
        assert out.shape == ref.shape == (rows, cols)
        assert np.max(np.abs(out - ref)) < 1e-7


def test_c_port_matches_numpy_oracle():
    from tests._util import make_params, smooth_field

    for (X, Y, res) in [(40, 40, 1.0), (33, 21, 1.5), (10, 10, 4.0)]:
        cfg = oracle_cfg(make_params(X, Y, res, 8, 20, 6, kappa=0.3, thr=0.5))
        ccfg = c_oracle.make_cfg(X, Y, res, value_threshold=0.5, interval_factor=0.3)
        rng = np.random.RandomState(1)
        B = 12
        gt = np.stack([smooth_field(rng, (Y, X)) for _ in range(B)])
        mean, var = rng.uniform(0, 1, (B, Y, X)), rng.uniform(0.05, 2, (B, Y, X))
        st = orc.BatchState(gt=gt.copy(), mean=mean.copy(), var=var.copy(), prev=np.tile([2.0, 2.0, 14.0], (B, 1)))
        cm, cv, cp = mean.copy(), var.copy(), st.prev.copy()
        for t in range(4):
            a = np.ascontiguousarray(np.stack([rng.uniform(0, X * res, B), rng.uniform(0, Y * res, B), rng.uniform(5, 20, B)], axis=1))
            adaptive, mode = t % 2 == 1, t % 2
            if t < 2:
                ro = orc.batched_full_step(cfg, st, a, seed=99, env_offset=7, adaptive=adaptive, reward_mode=mode)
                rc = c_oracle.step(ccfg, gt, cm, cv, cp, a, None, seed=99, env_offset=7, step_idx=t, flags=mode | (4 if adaptive else 0))
            else:
                eps = rng.standard_normal((B, X * Y))
                ro = orc.batched_full_step(cfg, st, a, eps=eps, adaptive=adaptive, reward_mode=mode)
                rc = c_oracle.step(ccfg, gt, cm, cv, cp, a, eps, flags=mode | (4 if adaptive else 0))
            assert np.max(np.abs(ro - rc)) < 1e-11 and np.max(np.abs(cm - st.mean)) < 1e-13 and np.max(np.abs(cv - st.var)) < 1e-13


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, out in kat:
        got = orc.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert tuple(int(x) for x in got) == out
    n = orc.device_normals(1, 2, 3, 50000)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01


# ---- MCTS-zero rollout loop: oracle/mcts_oracle.py vs the REAL reference MCTS (tests/golden/make_golden_mcts.py) ----------
@pytest.mark.parametrize("case", ["A", "B", "C"])
@pytest.mark.parametrize("deploy", [False, True])
def test_mcts_oracle_matches_reference_search(case, deploy):
    import json

    from oracle import mcts_oracle as morc
    from tests._util import stub_policy_value

    g = golden("golden_mcts.npz")
    params = params_from_json(g[f"{case}_cfg"])
    cfg = oracle_cfg(params)
    hyper, meta = json.loads(str(g[f"{case}_hyper"])), json.loads(str(g[f"{case}_meta"]))
    budget, sims = meta["budget"], hyper["num_mcts_simulations"]
    num_actions = 50
    tag = f"{case}_{'deploy' if deploy else 'train'}"

    def ev(info):
        return stub_policy_value(info["previous_action"], info["budget"] / budget, num_actions)

    o = morc.OracleMCTS(cfg, hyper, meta["episode_horizon"], evaluator=ev)
    o.search(g[f"{case}_var0"].copy(), g[f"{case}_prev"].copy(), budget, sims, root_noise=g[f"{case}_noise"])
    assert np.array_equal(o.Nsa[()], g[f"{tag}_Nsa"])  # visit counts: exact
    assert np.array_equal(o.Vs[()], g[f"{tag}_Vs"])
    assert o.Ns[()] == int(g[f"{tag}_Ns"]) and o.inference_counter == int(g[f"{tag}_inferences"])
    assert np.max(np.abs(o.Qsa[()] - g[f"{tag}_Qsa"])) <= 1e-12
    assert np.max(np.abs(o.Ps[()] - g[f"{tag}_Ps"])) <= 1e-15
    pol, _ = o.policy_from_root(temperature=1, deploy_time=deploy)
    assert np.max(np.abs(pol - g[f"{tag}_policy"])) <= 1e-15


def test_host_ground_truth_generator_matches_reference():
    """ipp_rl_b200/simulations/ground_truths.py (host, vectorised) vs the reference's gaussian_random_field under the same seed."""
    from ipp_rl_b200.simulations import ground_truths as gtgen

    g = golden("golden_grf.npz")
    for name in g["names"]:
        X, Y = (int(v) for v in g[f"{name}_dims"])
        r = float(g[f"{name}_radius"])
        np.random.seed(int(g[f"{name}_seed"]))
        f = gtgen.gaussian_random_field(lambda k: k ** (-r), X, Y)
        assert f.shape == (Y, X)
        assert np.max(np.abs(f - g[f"{name}_field"])) <= 1e-12


# --------------------------------------------------------------------------------------------------
# experience path: value targets, prioritised replay, shift augmentation
# --------------------------------------------------------------------------------------------------
def test_experience_oracle_matches_reference():
    from oracle import experience_oracle as xo

    g = golden("golden_experience.npz")
    for k in g["vt_cases"]:
        gamma, H = g[f"vt{k}_params"]
        vals, total = xo.value_targets(g[f"vt{k}_rewards"], float(gamma), int(H))
        assert np.max(np.abs(vals - g[f"vt{k}_values"])) <= 1e-14
        assert abs(total - float(g[f"vt{k}_total"])) <= 1e-13
    alpha = float(g["per_alpha"])
    for t in range(int(g["per_rounds"])):
        idx, w = xo.prioritized_sample(g[f"per{t}_priorities"], alpha, float(g[f"per{t}_beta"]), g[f"per{t}_uniforms"])
        assert np.array_equal(idx, g[f"per{t}_indices"])
        assert np.array_equal(w, g[f"per{t}_weights"])
        assert np.array_equal(g["per_states"][idx], g[f"per{t}_states"])
    base = g["per_states"][g["aug_sel"]]
    expect = np.vstack([base] + [xo.shift_with_replication(base, int(dy), int(dx)) for dy, dx in g["aug_offsets"]])
    assert np.array_equal(expect, g["aug_states"])
