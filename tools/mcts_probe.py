"""The tree-search leg of bench.py alone (16 384 trees, 200 x 200 grid, horizon 5, uniform priors): ms per lock-step simulation.
    IPP_B200_LIB=build/variants/libipp_X.so python tools/mcts_probe.py [layout] [sims]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipp_rl_b200 import BatchedEngine, EngineConfig, _capi as capi
from ipp_rl_b200.planning.mcts_zero import BatchedMCTS

T = 16384
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
W = dict(x_dim=200, y_dim=200, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0)
stream = torch.cuda.Stream()
eng = BatchedEngine(EngineConfig(batch=T, layout=capi.LAYOUT_NAMES[sys.argv[1] if len(sys.argv) > 1 else "split"], seed=20260925, stream=stream.cuda_stream, **W))
eng.reset(0.5, 1.82)
eng.synth_ground_truth(1000)
rng = np.random.RandomState(777)
for t in range(4):
    eng.step(rng.randint(0, eng.num_actions, T).astype(np.int32))
hyper = dict(puct_init=15.0, puct_base=10000, num_mcts_simulations=S, gamma=1.0, dirichlet_alpha=0.3, dirichlet_eps=0.25, forced_playout_factor=2.0,
             max_valid_action_distance=11.5)
budgets = np.full(T, 150.0, np.float32)
out = []
with torch.cuda.stream(stream):
    with BatchedMCTS(eng, hyper, dict(episode_horizon=5, scenario_info=None), n_trees=T) as mcts:
        for rep in range(3):
            mcts.begin(budgets)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(S):
                mcts.simulate(None)
            e1.record(stream)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / S)
        st = mcts.root_stats()
        chk = int((st["Nsa"].astype(np.int64) * (np.arange(st["Nsa"].shape[1]) + 1)).sum() % 1000003)
print(os.environ.get("IPP_B200_LIB", "default"), "ms/sim:", " ".join(f"{x:.4f}" for x in out), "checksum", chk)
