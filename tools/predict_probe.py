"""Which side of the memory system bounds the persistent kernel?  Whole-batch prediction steps (light arithmetic, same staged bytes as a
full step) with and without the write-back, against the full step, on the C3 workload."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

B = 65536
W = dict(x_dim=200, y_dim=200, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0)
stream = torch.cuda.Stream()
LAYOUT = sys.argv[1] if len(sys.argv) > 1 else "super"
eng = BatchedEngine(EngineConfig(batch=B, layout=capi.LAYOUT_NAMES[LAYOUT], seed=1, stream=stream.cuda_stream, **W))
print(f"layout {LAYOUT}")
eng.reset(0.5, 1.82)
eng.synth_ground_truth(1)
rng = np.random.RandomState(0)
ids = torch.from_numpy(rng.randint(0, eng.num_actions, size=(32, B)).astype(np.int32)).cuda()
r = torch.empty(B, dtype=torch.float32, device="cuda")


def run(name, fn, n=100):
    with torch.cuda.stream(stream):
        for t in range(10):
            fn(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(n):
            fn(t)
        e1.record(stream)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:28s} {us:7.1f} us/launch  {B / us:7.1f} M/s")


run("full step (trace)", lambda t: eng.step_device(action_ids_ptr=ids[t % 32].data_ptr(), reward_ptr=r.data_ptr()))
run("predict, commit", lambda t: eng.predict_device(B, action_ids_ptr=ids[t % 32].data_ptr(), reward_ptr=r.data_ptr(), commit=True))
run("predict, no commit (no stores)", lambda t: eng.predict_device(B, action_ids_ptr=ids[t % 32].data_ptr(), reward_ptr=r.data_ptr(), commit=False))
run("full step (trace)", lambda t: eng.step_device(action_ids_ptr=ids[t % 32].data_ptr(), reward_ptr=r.data_ptr()))
