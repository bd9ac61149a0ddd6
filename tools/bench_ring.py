#!/usr/bin/env python
"""Experience-ring micro-benchmark (GPU): gather of training batches of observation planes out of the ring in HBM,
with and without the shift augmentation, against the HBM roofline (algorithmic bytes = 4 B read + 4 B written per
plane element), plus the sampling chain's latency.   usage: python tools/bench_ring.py [capacity=4096] [batch=256]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ipp_rl_b200 import _capi as capi  # noqa: E402
from ipp_rl_b200.planning.mcts_zero.replay_buffers import ExperienceRing  # noqa: E402

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
shape = (6, 200, 200)
P = 1875
peak = 6451.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
stream = torch.cuda.Stream()
ring = ExperienceRing(cap, shape, P, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    src = torch.rand((256,) + shape, device="cuda")
    val = torch.rand(256, device="cuda")
    for k in range(cap // 256):
        ring.push_device(256, src.data_ptr(), val.data_ptr(), val.data_ptr())
    out = torch.empty((n,) + shape, device="cuda")
    sh = torch.randint(-4, 5, (n, 2), dtype=torch.int8, device="cuda")
    res = {}
    for name, shp in (("gather", 0), ("gather+shift", sh.data_ptr())):
        for it in range(3):
            ring.sample_indices(n, alpha=0.75, beta=0.5, seed=it)
            ring.gather_device(n, obs_ptr=out.data_ptr(), shifts_ptr=shp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 20
        stream.synchronize()
        e0.record(stream)
        for it in range(K):
            ring.gather_device(n, obs_ptr=out.data_ptr(), shifts_ptr=shp)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / K
        gb = 8.0 * n * np.prod(shape) / 1e9
        res[name] = {"ms": ms, "GB/s": gb / (ms * 1e-3), "frac_of_hbm_peak": gb / (ms * 1e-3) / peak}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for it in range(50):
        ring._ck(ring._lib.ipp_ring_sample(ring._h, n, 0.75, 0.5, None, it, None, None, 1))
    e1.record(stream)
    stream.synchronize()
    res["sample_chain_us"] = e0.elapsed_time(e1) / 50 * 1e3
print(json.dumps({"ring": {"capacity": cap, "obs": shape, "batch": n, "bytes": ring.device_bytes}, "peak_GB/s": peak, **res}))
