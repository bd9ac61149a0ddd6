import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._util import engine_cfg, make_params, smooth_field
from ipp_rl_b200 import BatchedEngine

X, Y, res, a0, a1, da = 64, 64, 2.0, 6, 30, 8
params = make_params(X, Y, res, a0, a1, da, kappa=0.25, thr=0.45)
B, T = 96, 4
rng = np.random.RandomState(3)
gt = np.stack([smooth_field(rng, (Y, X)) for _ in range(8)])[np.arange(B) % 8]
mean0 = rng.uniform(0, 1, (B, Y, X)).astype(np.float32)
var0 = rng.uniform(0.05, 2.0, (B, Y, X)).astype(np.float32)
out = {}
for layout, path in [(1, "lsu"), (3, "lsu"), (3, "async")]:
    with BatchedEngine(engine_cfg(params, B, layout=layout, seed=99)) as eng:
        eng.set_step_path(path)
        eng.reset(0.5, 1.82)
        eng.set_ground_truth(gt)
        eng.set_state(mean0, var0)
        g0 = eng.get_ground_truth(); m0, v0 = eng.get_state()
        print(layout, path, "roundtrip gt", np.array_equal(g0, gt.astype(np.float32)), "state", np.array_equal(m0, mean0), np.array_equal(v0, var0))
        r2 = np.random.RandomState(17)
        res_ = []
        for t in range(T):
            ids = r2.randint(0, eng.num_actions, B).astype(np.int32)
            if t == 1:
                ids[:8] = [0, X - 1, X * (Y - 1), X * Y - 1, eng.num_actions - 1, eng.num_actions - X, X * Y, 2 * X * Y - 1][:8]
            if t == 2:
                noise = r2.standard_normal((B, eng.max_measurements)).astype(np.float32)
                r, z = eng.step(ids, noise=noise, reward_mode=0, adaptive=False, return_measurements=True)
            else:
                r = eng.step(ids, reward_mode=0, adaptive=False)
                z = None
            m, v = eng.get_state()
            res_.append((ids.copy(), r.copy(), m, v, z))
        out[(layout, path)] = res_
ref = out[(1, "lsu")]
N = X * Y
for key in [(3, "lsu"), (3, "async")]:
    for t in range(T):
        ids, r, m, v, z = out[key][t]
        ids0, r0, m00, v00, z0 = ref[t]
        bad_r = np.flatnonzero(r != r0)
        bad_m = np.flatnonzero((m != m00).reshape(B, -1).any(1))
        bad_v = np.flatnonzero((v != v00).reshape(B, -1).any(1))
        print(key, "t", t, "bad rewards", len(bad_r), "bad mean envs", len(bad_m), "bad var envs", len(bad_v), "z", None if z is None else int((z != z0).sum()))
        for b in bad_r[:6]:
            lvl = ids[b] // N; i = ids[b] % N; col, row = i // X, i % X
            print("   env", b, "id", ids[b], "lvl", lvl, "col", col, "row", row, "r", r[b], r0[b], "dm cells", int((m[b] != m00[b]).sum()), "dv cells", int((v[b] != v00[b]).sum()),
                  "maxdm", float(np.abs(m[b] - m00[b]).max()), "maxdv", float(np.abs(v[b] - v00[b]).max()))
            if (v[b] != v00[b]).any():
                rr, cc = np.nonzero(v[b] != v00[b]); print("      var diff rows", rr.min(), rr.max(), "cols", cc.min(), cc.max())
