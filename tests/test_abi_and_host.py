"""CPU: the C-ABI library loads and exports every symbol include/ipp_b200.h declares (no compute calls
without a GPU), the ctypes structs match the header's layout, the host-side mirror of the reference
interface behaves like the reference (names, validation, error behaviour), and the env-batch sharding
works across 2 processes (gloo)."""
import ctypes
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ipp_b200.h")
HEADERS = [HEADER, os.path.join(ROOT, "include", "ipp_mcts.h"), os.path.join(ROOT, "include", "ipp_experience.h")]


def _declared_functions():
    out = set()
    for h in HEADERS:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        out |= set(re.findall(r"\b(ipp_[a-z_0-9]+)\s*\(", src))
    return sorted(out)


def test_library_exports_every_declared_symbol():
    from ipp_rl_b200 import _capi

    lib = _capi.load_library()
    declared = _declared_functions()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert sorted(_capi.SIGNATURES) == declared, "ctypes SIGNATURES must list exactly the header's functions"


def test_ctypes_structs_match_the_header(tmp_path):
    from ipp_rl_b200 import _capi

    probe = tmp_path / "sz.c"
    probe.write_text(
        f'#include "{HEADERS[1]}"\n#include "{HEADERS[2]}"\n#include <stdio.h>\n#include <stddef.h>\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(ipp_config), sizeof(ipp_info), offsetof(ipp_config, resolution), "
        "offsetof(ipp_config, seed), offsetof(ipp_info, altitude), sizeof(ipp_mcts_config), sizeof(ipp_mcts_info), "
        "offsetof(ipp_mcts_config, puct_init), offsetof(ipp_mcts_info, device_bytes), sizeof(ipp_ring_config), sizeof(ipp_ring_info), "
        "offsetof(ipp_ring_config, stream), offsetof(ipp_ring_info, pushed)); return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(probe), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_capi.ipp_config), ctypes.sizeof(_capi.ipp_info), _capi.ipp_config.resolution.offset,
                     _capi.ipp_config.seed.offset, _capi.ipp_info.altitude.offset, ctypes.sizeof(_capi.ipp_mcts_config),
                     ctypes.sizeof(_capi.ipp_mcts_info), _capi.ipp_mcts_config.puct_init.offset, _capi.ipp_mcts_info.device_bytes.offset,
                     ctypes.sizeof(_capi.ipp_ring_config), ctypes.sizeof(_capi.ipp_ring_info), _capi.ipp_ring_config.stream.offset,
                     _capi.ipp_ring_info.pushed.offset]


def test_header_constants_match_the_ctypes_module():
    """Every numeric #define of the headers that _capi mirrors (layouts, flags, options, zero-copy bits, pointers) has the same
    value on both sides — a renumbered option would otherwise go unnoticed until a GPU run."""
    import re

    from ipp_rl_b200 import _capi

    defines = {}
    for h in HEADERS:
        for m in re.finditer(r"^#define\s+(IPP_[A-Z0-9_]+)\s+(-?\d+)u?\b", open(h).read(), re.M):
            defines[m.group(1)] = int(m.group(2))
    assert defines["IPP_LAYOUT_SPLIT"] == 4 and len(defines) > 30
    checked = 0
    for name, value in vars(_capi).items():
        if not name.isupper() or not isinstance(value, int) or isinstance(value, bool):
            continue
        for cname in (name, "IPP_" + name):
            if cname in defines:
                assert defines[cname] == value, (name, cname, value, defines[cname])
                checked += 1
                break
    assert checked >= 50, checked
    assert _capi.LAYOUT_NAMES == {"planes": 0, "mv": 1, "tiled": 2, "super": 3, "split": 4}


def test_missing_library_fails_loudly(tmp_path):
    from ipp_rl_b200 import _capi

    with pytest.raises(_capi.IppLibraryError):
        _capi.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ipp_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "ipp_oracle" not in txt or f.endswith((".cuh", ".cu")) and "Mirrored in oracle" in txt, f


# ---- host mirror of the reference interface ---------------------------------------------------------
def _params():
    from tests._util import make_params

    return make_params(10, 10, 4, 8, 14, 6)


def test_factories_validate_like_the_reference():
    from ipp_rl_b200.mapping.grid_maps import GridMap
    from ipp_rl_b200.sensors.models.sensor_model_factories import SensorModelFactory
    from ipp_rl_b200.sensors.models.sensor_models import AltitudeSensorModel
    from ipp_rl_b200.sensors.sensor_factories import SensorFactory
    from ipp_rl_b200.sensors.cameras import RGBCamera

    p = _params()
    gm = GridMap(p)
    assert (gm.x_dim, gm.y_dim, gm.resolution, gm.num_grid_cells) == (10, 10, 4, 100)
    model = SensorModelFactory(p).create_sensor_model()
    assert isinstance(model, AltitudeSensorModel) and (model.coeff_a, model.coeff_b) == (0.05, 0.2)
    sensor = SensorFactory(p, model, gm).create_sensor()
    assert isinstance(sensor, RGBCamera) and sensor.sensor_simulation is None
    # missing / unknown keys -> ValueError (reference: logger.error + bare raise ValueError)
    with pytest.raises(ValueError):
        GridMap({}).x_dim
    bad = _params()
    bad["sensor"]["model"]["type"] = "unknown_model"
    with pytest.raises(ValueError):
        SensorModelFactory(bad)
    bad = _params()
    del bad["sensor"]["model"]["coeff_b"]
    with pytest.raises(ValueError):
        SensorModelFactory(bad)
    bad = _params()
    bad["sensor"]["type"] = "lidar"
    with pytest.raises(ValueError):
        SensorFactory(bad, model, gm)
    bad = _params()
    del bad["sensor"]["encoding"]
    with pytest.raises(ValueError):
        SensorFactory(bad, model, gm)


def test_host_sensor_and_actions_match_reference_vectors():
    from ipp_rl_b200.mapping.grid_maps import GridMap
    from ipp_rl_b200.planning.common import actions
    from ipp_rl_b200.sensors.models.sensor_model_factories import SensorModelFactory
    from ipp_rl_b200.sensors.sensor_factories import SensorFactory
    from tests._util import golden, params_from_json

    g = golden("golden_sensor_actions_metrics.npz")
    for si in range(int(g["fp_count"])):
        p = params_from_json(g[f"fp{si}_cfg"])
        p["sensor"].update(type="rgb_camera", encoding="rgb8")
        p["sensor"]["model"]["type"] = "altitude_dependent"
        gm = GridMap(p)
        model = SensorModelFactory(p).create_sensor_model()
        s = SensorFactory(p, model, gm).create_sensor()
        for q, f, r, s2 in list(zip(g[f"fp{si}_poses"], g[f"fp{si}_fov"], g[f"fp{si}_rf"], g[f"fp{si}_sigma2"]))[::3]:
            assert s.project_field_of_view(q) == tuple(int(t) for t in f)
            assert s.get_resolution_factor(q) == r
            assert model.get_noise_variance(q) == pytest.approx(s2, abs=1e-16)
    gm = GridMap(_params())
    assert np.array_equal(actions.action_table(gm, 8, 14, 6), g["act_ex10_table"])
    assert np.array_equal(actions.action_dict_to_np_array(actions.enumerate_actions(gm, 8, 14, 6)), g["act_ex10_table"])
    uav = {"max_v": 2, "max_a": 2}
    for a, b, ft, ed in zip(g["cost_a"], g["cost_b"], g["cost_flight_time"], g["cost_distance"]):
        assert actions.action_costs(a, b, uav) == pytest.approx(ft, abs=1e-13)
        assert actions.action_costs(a, b, None) == pytest.approx(ed, abs=1e-13)
    acts = actions.get_actions(np.array([2.0, 2.0, 14.0]), 12.0, gm, 8, 14, 6, uav)
    want = [a for a in g["act_ex10_table"] if 0 < actions.action_costs(a, np.array([2.0, 2.0, 14.0]), uav) <= 12.0]
    assert sorted(map(tuple, acts)) == sorted(map(tuple, want)) and len(acts) > 0


def test_host_metrics_and_rewards_match_reference_vectors():
    from ipp_rl_b200.mapping.grid_maps import DiagonalCovariance
    from ipp_rl_b200.planning import evaluation_metrics as em
    from ipp_rl_b200.planning.common import rewards
    from tests._util import golden

    g = golden("golden_sensor_actions_metrics.npz")
    for k, (Y, X) in enumerate(g["met_shapes"]):
        gt, mean, var, msk = (g[n][k, :Y, :X] for n in ("met_gt", "met_mean", "met_var", "met_mask"))
        cov, m = DiagonalCovariance(var), msk.astype(bool).ravel()
        with np.errstate(all="ignore"):
            got = [em.root_mean_squared_error(gt, mean), em.weighted_root_mean_squared_error(gt, mean), em.mean_log_loss(gt, mean, cov),
                   em.weighted_mean_log_loss(gt, mean, cov), em.map_uncertainty(cov), em.map_uncertainty_difference(cov, m),
                   em.root_mean_squared_error(gt, mean, m), em.map_uncertainty(cov, m)]
        for x, y in zip(g["met_values"][k], got):
            assert (np.isnan(x) and np.isnan(y)) or y == pytest.approx(x, rel=1e-12)
        assert np.array_equal(rewards.compute_adaptive_msk(mean, cov, 0.6, 0.3), m)
        assert np.array_equal(np.diag(np.asarray(cov)), var.ravel()) and np.trace(cov) == pytest.approx(var.sum())


def test_ground_truth_generators_are_seed_compatible():
    from ipp_rl_b200.simulations import ground_truths

    def ref_like(pk, x_dim, y_dim):  # straight transcription of the published algorithm, scalar loops
        noise = np.fft.fft2(np.random.normal(size=(y_dim, x_dim)))
        amp = np.zeros((y_dim, x_dim))
        for i, kx in enumerate(ground_truths.fft_indices(y_dim)):
            for j, ky in enumerate(ground_truths.fft_indices(x_dim)):
                amp[i, j] = 0.0 if (kx == 0 and ky == 0) else np.sqrt(pk(np.sqrt(float(kx) ** 2 + float(ky) ** 2)))
        f = np.fft.ifft2(noise * amp).real
        return (f - f.min()) / (f.max() - f.min())

    for x, y in [(10, 10), (17, 12), (32, 32)]:
        np.random.seed(5)
        a = ground_truths.gaussian_random_field(lambda k: k ** (-5.0), x, y)
        np.random.seed(5)
        b = ref_like(lambda k: k ** (-5.0), x, y)
        assert a.shape == (y, x) and np.max(np.abs(a - b)) < 1e-10 and a.min() == 0 and a.max() == 1


def test_facade_objects_pickle_without_cuda_handles():
    from ipp_rl_b200.mapping.grid_maps import DiagonalCovariance, GridMap

    gm = GridMap(_params())
    gm.mean, gm.cov_matrix = np.zeros((10, 10)), DiagonalCovariance(np.ones(100))
    setattr(gm, "_b200_backend", lambda: 0)  # stands for a live device backend: not picklable
    clone = pickle.loads(pickle.dumps(gm))
    assert not hasattr(clone, "_b200_backend") and clone.x_dim == 10 and np.trace(clone.cov_matrix) == 100


# ---- sharding (N > 1 path) -----------------------------------------------------------------------------
def test_shard_bounds_cover_the_batch():
    from ipp_rl_b200.distributed import all_shard_counts, shard_bounds, sharded_config
    from ipp_rl_b200 import EngineConfig

    for total, world in [(65536, 8), (131072, 8), (10, 3), (7, 7), (100, 1)]:
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert all_shard_counts(total, world) == [c for _, c in spans]
    cfg = sharded_config(EngineConfig(batch=1, env_id_offset=5), 10, 3, 2, device=2)
    assert (cfg.batch, cfg.env_id_offset, cfg.device) == (3, 5 + 7, 2)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from ipp_rl_b200.distributed import gather_experience, gather_rewards, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
for total in (10, 7):
    first, count = shard_bounds(total, 2, rank)
    local = torch.arange(first, first + count, dtype=torch.float32) * 1.5   # "reward" of global env i = 1.5 i
    out = gather_rewards(local, total)
    assert torch.equal(out, torch.arange(total, dtype=torch.float32) * 1.5), (rank, out)
    ids = torch.arange(first, first + count, dtype=torch.float32)
    exp = gather_experience({{"obs": ids[:, None, None].expand(count, 2, 3) + 0.25, "values": ids * 2}}, total)
    full = torch.arange(total, dtype=torch.float32)
    assert torch.equal(exp["obs"], full[:, None, None].expand(total, 2, 3) + 0.25) and torch.equal(exp["values"], full * 2)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_rewards_two_processes_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
        assert "ok" in so


class _OracleRing:
    """CPU stand-in for ExperienceRing (same methods, NumPy + oracle/experience_oracle.py): lets the host classes that mirror
    the reference's replay buffers be checked without a GPU.  The CUDA ring itself is checked in tests/test_gpu_experience.py."""

    def __init__(self, states, policies, values, rewards, masks):
        self.s, self.p, self.v, self.r, self.m = states, policies, values, rewards, masks
        self.pri = np.ones(len(states), np.float32)

    def __len__(self):
        return len(self.s)

    def reset_priorities(self):
        self.pri[:] = np.float32(1.0 / len(self.s))

    def priorities(self):
        return self.pri.copy()

    def update_priorities(self, idx, pr):
        self.pri[np.asarray(idx)] = np.asarray(pr, np.float32)

    def sample_indices(self, n, alpha=-1.0, beta=0.0, uniforms=None, seed=0):
        from oracle import experience_oracle as xo

        if alpha < 0:
            return xo.uniform_sample(len(self.s), uniforms), np.ones(n, np.float32)
        return xo.prioritized_sample(self.pri.astype(np.float64), alpha, beta, uniforms)

    def gather(self, indices=None, n=None, shifts=None, with_policy=True):
        from oracle import experience_oracle as xo

        idx = np.asarray(indices)
        obs = self.s[idx].copy()
        if shifts is not None:
            obs = np.stack([xo.shift_with_replication(o, int(dy), int(dx)) for o, (dy, dx) in zip(obs, np.asarray(shifts))])
        return obs, self.p[idx], self.m[idx].astype(np.uint8), self.v[idx], self.r[idx]


def test_replay_buffer_classes_follow_the_reference_sequence():
    """Host logic of ExperienceReplayBuffer / PrioritizedExperienceReplayBuffer (sample sizes, beta annealing, augmentation order,
    return tuple) against the sequence the REAL reference buffer produced (golden_experience.npz), on a CPU stand-in ring."""
    from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceReplayBuffer, PrioritizedExperienceReplayBuffer

    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_experience.npz"))

    class Scripted:
        def __init__(self, uniforms=(), ints=()):
            self.u, self.i = list(uniforms), list(ints)

        def random_sample(self, n):
            return self.u.pop(0)

        def randint(self, lo, hi, size):
            return np.asarray(self.i.pop(0))

    ring = _OracleRing(g["per_states"], g["per_policies"], g["per_values"], g["per_rewards"], g["per_masks"])
    rounds = int(g["per_rounds"])
    buf = PrioritizedExperienceReplayBuffer(ring, batch_size=32, alpha=float(g["per_alpha"]), beta0=0.5, num_epochs=3,
                                            rng=Scripted(uniforms=[g[f"per{t}_uniforms"] for t in range(rounds)]))
    assert buf.sample_size == 32 and buf.total_steps == (257 // 32) * 3
    for t in range(rounds):
        assert abs(float(buf.beta) - float(g[f"per{t}_beta"])) < 1e-12
        states, policies, values, rewards, msk, idx, w = buf.sample()
        assert np.array_equal(idx, g[f"per{t}_indices"]) and np.allclose(w, g[f"per{t}_weights"], rtol=2e-6)
        assert np.array_equal(states, g[f"per{t}_states"]) and msk.dtype == bool
        buf.update(idx, g[f"per{t}_new_priorities"])
        buf.step()
    assert np.allclose(buf.priorities, g["per_final_priorities"], rtol=1e-6)
    assert abs(float(buf.beta) - float(g["per_final_beta"])) < 1e-12

    sel = g["aug_sel"]
    u = (sel + 0.5) / len(ring)  # uniform draws that land on the reference's four samples
    ubuf = ExperienceReplayBuffer(ring, batch_size=16, num_augmented_samples=3, rng=Scripted(uniforms=[u], ints=[g["aug_offsets"] + 4]))
    assert ubuf.sample_size == 4
    states, policies, values, rewards, msk, idx, w = ubuf.sample()
    assert np.array_equal(idx, sel) and np.array_equal(states, g["aug_states"]) and np.all(w == 1)
    assert np.array_equal(values, g["aug_values"]) and np.array_equal(policies, g["aug_policies"])
