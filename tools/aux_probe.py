"""Throughput of the streaming kernels around the step (C3 grid, 4 096 envs, device-resident outputs): observation planes
(ipp_observe, f2), evaluation metrics (ipp_eval), prior reset, against their algorithmic bytes and the HBM copy peak."""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

peak, _ = bench.load_peaks()
B, X = 4096, 200
W = dict(x_dim=X, y_dim=X, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0)
stream = torch.cuda.Stream()
for layout in ("super", "split", "planes"):
    eng = BatchedEngine(EngineConfig(batch=B, layout=capi.LAYOUT_NAMES[layout], seed=1, stream=stream.cuda_stream, **W))
    eng.reset(0.5, 1.82)
    eng.synth_ground_truth(1)
    out = torch.empty((B, 6, X, X), dtype=torch.float32, device="cuda")
    met = torch.empty((B, capi.NUM_METRICS), dtype=torch.float32, device="cuda")
    lib = eng._lib

    def timed(fn, n=10):
        with torch.cuda.stream(stream):
            fn(); fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    cells = B * X * X
    ms = timed(lambda: eng._ck(lib.ipp_observe(eng._h, 0, B, None, None, capi.OBS_COSTS, C.c_void_p(out.data_ptr()), 1)))
    alg = cells * (8 + 6 * 4)  # read mean + var, write six planes
    print(f"{layout:7s} observe  {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s algorithmic  frac {alg / ms / 1e6 / peak:5.2f}")
    ms = timed(lambda: eng._ck(lib.ipp_eval_device(eng._h, C.c_void_p(met.data_ptr()))))
    alg = cells * 12 * 2  # two passes over gt, mean, var
    print(f"{layout:7s} eval     {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s algorithmic  frac {alg / ms / 1e6 / peak:5.2f}")
    ms = timed(lambda: eng.reset(0.5, 1.82))
    alg = cells * 8
    print(f"{layout:7s} reset    {ms:7.3f} ms  {alg / ms / 1e6:7.1f} GB/s algorithmic  frac {alg / ms / 1e6 / peak:5.2f}")
    eng.close()
