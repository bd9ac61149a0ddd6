"""CPU baseline for bench.py: the oracle's C port of the reference algorithm on all host cores.

TEST INFRASTRUCTURE.  kind = "port": the reference is pure Python and cannot travel to the GPU box
(and its dense N x N covariance is infeasible at 200x200: 12.8 GB per env), so the
like-for-like CPU number is this float64 restatement of the same per-cell algorithm (pinned against
the real reference through tests/golden), OpenMP-parallel over envs.
"""
import time

import numpy as np

from . import c_oracle


def _synthetic_gt(rng, n_maps, Y, X):
    yy, xx = np.mgrid[0:Y, 0:X].astype(np.float64)
    out = np.empty((n_maps, Y, X))
    for i in range(n_maps):
        f = np.zeros((Y, X))
        for _ in range(5):
            kx, ky = rng.uniform(-0.3, 0.3, 2)
            f += rng.uniform(0.3, 1.0) * np.sin(kx * xx + ky * yy + rng.uniform(0, 6.28))
        out[i] = (f - f.min()) / (f.max() - f.min())
    return out


def run(workload: dict, steps: int = 3, warmup: int = 1, envs: int = 4096, reward_mode: int = 1, min_seconds: float = 0.0) -> dict:
    X, Y, res = workload["x_dim"], workload["y_dim"], workload["resolution"]
    cfg = c_oracle.make_cfg(X, Y, res, workload.get("angle_x", 60.0), workload.get("angle_y", 60.0), workload.get("coeff_a", 0.05),
                            workload.get("coeff_b", 0.2), 10.0, workload.get("max_v", 2.0), workload.get("max_a", 2.0))
    build_note = c_oracle.use_native_build()
    c_oracle.use_all_cores()
    rng = np.random.RandomState(4242)
    base = _synthetic_gt(rng, 32, Y, X)
    gt = np.ascontiguousarray(base[np.arange(envs) % 32])
    mean = np.full((envs, Y, X), 0.5)
    var = np.full((envs, Y, X), 1.82)
    prev = np.tile([2.0, 2.0, 14.0], (envs, 1))
    n_lv = int((workload["max_altitude"] - workload["min_altitude"]) / workload["altitude_spacing"]) + 1
    alts = np.linspace(workload["min_altitude"], workload["max_altitude"], n_lv)

    def actions():
        col = rng.randint(0, X, envs)
        row = rng.randint(0, Y, envs)
        return np.ascontiguousarray(np.stack([res * col + 0.5 * res, res * row + 0.5 * res, alts[rng.randint(0, n_lv, envs)]], axis=1))

    flags = reward_mode & 3
    acts = [actions() for _ in range(warmup + steps)]
    for t in range(warmup):
        c_oracle.step(cfg, gt, mean, var, prev, acts[t], None, seed=20260925, step_idx=t, flags=flags)
    t0 = time.perf_counter()
    done = 0
    t = warmup
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        c_oracle.step(cfg, gt, mean, var, prev, acts[warmup + (done % steps)], None, seed=20260925, step_idx=t, flags=flags)
        done += 1
        t += 1
    dt = time.perf_counter() - t0
    return {
        "steps_per_sec": envs * done / dt,
        "ms_per_step": 1e3 * dt / done,
        "cores": c_oracle.num_threads(),
        "kind": "port",
        "envs": envs,
        "sample": f"{done} steps x {envs} envs of the same 200x200 / 3-altitude workload, fp64 C port (oracle/ipp_oracle.c), OpenMP "
                  f"over envs, Philox noise, {build_note}; {dt:.2f} s wall",
    }
