#!/usr/bin/env python
"""Throughput of the fused step on the OTHER BASELINE.json configurations (the headline C3 line is bench.py's):
C2 (4 096 envs, 50x50 @ 4 m, fixed altitude 8 m / 14 m), C4-style predict-only steps (16 384 envs, 200x200), C5 (4 096 envs/GPU,
400x400, variance-reduction reward), next to C3 — device-resident inputs, CUDA events on the engine's stream, algorithmic
bytes = (20 | 8) B x covered cells + 16 B per env-step.  One JSON line per config.   usage: python tools/bench_configs.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (footprint_cells, load_peaks)
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi  # noqa: E402

peak, peak_src = bench.load_peaks()
CONFIGS = [
    # name, batch, grid, res, altitudes (min, max, spacing), levels used, reward mode, predict-only
    ("C2 50x50 res4 h=8 (rf1, 3x3)", 4096, 50, 4.0, (8.0, 14.0, 6.0), [0], capi.REWARD_TRACE, False),
    ("C2 50x50 res4 h=14 (rf2, 5x5)", 4096, 50, 4.0, (8.0, 14.0, 6.0), [1], capi.REWARD_TRACE, False),
    ("C3 200x200 res1 3 altitudes, entropy reward", 65536, 200, 1.0, (8.0, 20.0, 6.0), [0, 1, 2], capi.REWARD_GAUSS_ENTROPY, False),
    ("C4-style predict-only steps 200x200 (16384 envs)", 16384, 200, 1.0, (8.0, 20.0, 6.0), [0, 1, 2], capi.REWARD_TRACE, True),
    ("C5 400x400 res1 3 altitudes, variance-reduction reward (4096 envs)", 4096, 400, 1.0, (8.0, 20.0, 6.0), [0, 1, 2], capi.REWARD_TRACE, False),
    ("C5 at 32768 envs on one GPU (62.9 GB)", 32768, 400, 1.0, (8.0, 20.0, 6.0), [0, 1, 2], capi.REWARD_TRACE, False),
]
K, W, POOL = 300, 10, 16
stream = torch.cuda.Stream()
for name, B, G, res, (amin, amax, asp), levels, mode, predict in CONFIGS:
    # stepping engines on super-tiles, covariance-only (search) engines on the split layout
    layout = capi.LAYOUT_SPLIT if predict else capi.LAYOUT_SUPER
    cfg = EngineConfig(batch=B, x_dim=G, y_dim=G, resolution=res, min_altitude=amin, max_altitude=amax, altitude_spacing=asp,
                       layout=layout, seed=1, stream=stream.cuda_stream)
    with BatchedEngine(cfg) as eng, torch.cuda.stream(stream):
        eng.reset(0.5, 1.82)
        eng.synth_ground_truth(seed=3)
        info = eng.info
        radii = [info.radius_x[k] for k in range(info.num_altitude_levels)]
        rng = np.random.RandomState(0)
        N = G * G
        ids = (rng.choice(levels, size=(POOL, B)) * N + rng.randint(0, N, size=(POOL, B))).astype(np.int32)
        cells = bench.footprint_cells(ids.astype(np.int64), G, G, radii).sum(axis=1)
        ids_dev = torch.from_numpy(ids).cuda()
        rew = torch.empty(B, dtype=torch.float32, device="cuda")

        def go(t):
            if predict:
                eng.predict_device(B, action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=rew.data_ptr(), commit=True, reward_mode=mode)
            else:
                eng.step_device(action_ids_ptr=ids_dev[t % POOL].data_ptr(), reward_ptr=rew.data_ptr(), reward_mode=mode)

        la0, ll0 = eng.path_launches("async"), eng.path_launches("lsu")
        for t in range(W):
            go(t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        e0.record(stream)
        for t in range(W, W + K):
            go(t)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / K
        per_cell = 8.0 if predict else 20.0
        alg = (per_cell * float(sum(cells[t % POOL] for t in range(W, W + K))) / K + 16.0 * B)
        gbs = alg / (ms * 1e-3) / 1e9
        print(json.dumps({"config": name, "batch": B, "ms_per_step": ms, "env_steps_per_sec": B / (ms * 1e-3), "algorithmic_GB/s": gbs,
                          "frac_of_hbm_peak": gbs / peak, "mean_cells": float(cells.mean() / B),
                          # the kernel that actually ran (launch counters), not the engine's path option
                          "kernel": ("ipp_step_bulk_kernel" + ("<MODE_PREDICT>" if predict else "")) if eng.path_launches("async") - la0 >= K
                          else ("ipp_step_kernel" + ("<MODE_PREDICT>" if predict else "")) if eng.path_launches("lsu") - ll0 >= K else "mixed",
                          "layout": "split" if predict else "super", "state_bytes": eng.device_bytes}))
