"""Per-step host latency of the synchronous e2e call (BatchedEngine.step with pinned host buffers), first 40 steps after an idle gap."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi

B = 65536
W = dict(x_dim=200, y_dim=200, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0)
stream = torch.cuda.Stream()
layout = capi.LAYOUT_NAMES[sys.argv[1] if len(sys.argv) > 1 else "super"]
eng = BatchedEngine(EngineConfig(batch=B, layout=layout, seed=1, stream=stream.cuda_stream, **W))
eng.reset(0.5, 1.82)
eng.synth_ground_truth(1)
rng = np.random.RandomState(0)
ids = torch.from_numpy(rng.randint(0, eng.num_actions, size=(64, B)).astype(np.int32)).pin_memory()
out = torch.empty(B, dtype=torch.float32).pin_memory()
ids_np, out_np = ids.numpy(), out.numpy()
rows = [ids_np[k] for k in range(64)]
for rep in range(2):
    time.sleep(0.5)
    ts = []
    for t in range(40):
        t0 = time.perf_counter()
        eng.step(rows[t], reward_mode=capi.REWARD_TRACE, out=out_np)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("rep", rep, "per-step us:", " ".join(f"{x:.0f}" for x in ts))
# device-side only, same ids already resident
d_ids = ids.cuda()
d_r = torch.empty(B, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    for t in range(40):
        eng.step_device(action_ids_ptr=d_ids[t].data_ptr(), reward_ptr=d_r.data_ptr(), reward_mode=capi.REWARD_TRACE)
    e1.record(stream)
torch.cuda.synchronize()
print("device-only us/step:", e0.elapsed_time(e1) * 1e3 / 40)
