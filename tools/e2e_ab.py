"""A/B of the synchronous host step's transfer modes inside ONE process (box-to-box and run-to-run spread is larger than the
differences): modes interleaved, several repetitions, median of the per-repetition means."""
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
MODES = {"copy ids": (False, False), "fetch ids": (True, False)}
res = {k: [] for k in MODES}
N, REPS = 300, 5
with torch.cuda.stream(stream):
    for rep in range(REPS):
        for name, (fetch, poll) in MODES.items():
            eng.set_zero_copy(rewards=True, ids=False, ids_fetch=fetch)
            for t in range(10):
                eng.step(rows[t], reward_mode=capi.REWARD_TRACE, out=out_np)
            t0 = time.perf_counter()
            for t in range(N):
                eng.step(rows[t % 64], reward_mode=capi.REWARD_TRACE, out=out_np)
            res[name].append((time.perf_counter() - t0) / N * 1e6)
for name, v in res.items():
    print(f"{name:18s} median {np.median(v):6.1f} us/step   reps: " + " ".join(f"{x:.1f}" for x in v))
