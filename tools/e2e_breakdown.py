"""Where do the microseconds of a synchronous host step go?  Device-side intervals (CUDA events around the H2D copy of the ids and the
fused kernel) beside the host wall clock of the same step, and the public call (BatchedEngine.step) for reference."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

B = 65536
W = dict(x_dim=200, y_dim=200, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0)
stream = torch.cuda.Stream()
eng = BatchedEngine(EngineConfig(batch=B, layout=capi.LAYOUT_SUPER, seed=1, stream=stream.cuda_stream, **W))
eng.reset(0.5, 1.82)
eng.synth_ground_truth(1)
rng = np.random.RandomState(0)
ids = torch.from_numpy(rng.randint(0, eng.num_actions, size=(64, B)).astype(np.int32)).pin_memory()
out = torch.empty(B, dtype=torch.float32).pin_memory()
ids_np, out_np = ids.numpy(), out.numpy()
rows = [ids_np[k] for k in range(64)]
d_ids = torch.empty(B, dtype=torch.int32, device="cuda")
d_r = torch.empty(B, dtype=torch.float32, device="cuda")
N = 200
with torch.cuda.stream(stream):
    for t in range(20):
        eng.step(rows[t], reward_mode=capi.REWARD_TRACE, out=out_np)
    # public call
    t0 = time.perf_counter()
    for t in range(N):
        eng.step(rows[t % 64], reward_mode=capi.REWARD_TRACE, out=out_np)
    pub = (time.perf_counter() - t0) / N * 1e6
    # the same sequence by hand with events
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(N)]
    host, submit = [], []
    for t in range(N):
        t0 = time.perf_counter()
        evs[t][0].record(stream)
        d_ids.copy_(ids[t % 64], non_blocking=True)
        evs[t][1].record(stream)
        eng.step_device(action_ids_ptr=d_ids.data_ptr(), reward_ptr=d_r.data_ptr(), reward_mode=capi.REWARD_TRACE)
        evs[t][2].record(stream)
        t1 = time.perf_counter()
        stream.synchronize()
        t2 = time.perf_counter()
        host.append((t2 - t0) * 1e6)
        submit.append((t1 - t0) * 1e6)
    copy = np.array([e[0].elapsed_time(e[1]) * 1e3 for e in evs])
    kern = np.array([e[1].elapsed_time(e[2]) * 1e3 for e in evs])
    # back-to-back kernels (no host in the loop)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(N):
        eng.step_device(action_ids_ptr=d_ids.data_ptr(), reward_ptr=d_r.data_ptr(), reward_mode=capi.REWARD_TRACE)
    e1.record(stream)
    stream.synchronize()
    b2b = e0.elapsed_time(e1) * 1e3 / N
med = lambda x: float(np.median(x))
print(f"public BatchedEngine.step          {pub:7.1f} us/step")
print(f"by hand: host wall                 {med(host):7.1f} us   (submit calls return after {med(submit):.1f} us)")
print(f"  device: H2D ids (events)         {med(copy):7.1f} us")
print(f"  device: kernel in the e2e loop   {med(kern):7.1f} us")
print(f"  host wall - device intervals     {med(np.array(host) - copy - kern):7.1f} us")
print(f"kernel back to back                {b2b:7.1f} us")
