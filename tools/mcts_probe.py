"""The tree-search leg of bench.py alone (16 384 trees, 200 x 200 grid, horizon 5): ms per lock-step simulation.
    IPP_B200_LIB=build/variants/libipp_X.so python tools/mcts_probe.py [layout] [sims] [uniform|peaked]
uniform: no evaluator (uniform priors, zero values) — the search exploits one path, few new edges; peaked: synthetic network
outputs in device memory (soft-max of random logits per tree, small random values) — the tree grows by about one edge and one
expansion per simulation, as under a trained policy / value network."""
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
MODE = sys.argv[3] if len(sys.argv) > 3 else "uniform"
out = []
with torch.cuda.stream(stream):
    with BatchedMCTS(eng, hyper, dict(episode_horizon=5, scenario_info=None), n_trees=T) as mcts:
        pri = val = None
        if MODE == "peaked":
            g = torch.Generator(device="cuda").manual_seed(5)
            pri = torch.softmax(4.0 * torch.randn(T, mcts.window_slots, device="cuda", generator=g), dim=1).contiguous()
            val = (0.05 * torch.rand(T, device="cuda", generator=g)).contiguous()
            torch.cuda.synchronize()
        for rep in range(3):
            mcts.begin(budgets)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(S):
                if pri is None:
                    mcts.simulate(None)
                else:
                    mcts.simulate_device(priors_window_ptr=pri.data_ptr(), values_ptr=val.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / S)
        st = mcts.root_stats()
        chk = int((st["Nsa"].astype(np.int64) * (np.arange(st["Nsa"].shape[1]) + 1)).sum() % 1000003)
        edges = int(mcts.info.edges)
print(MODE, "edges per tree", round(edges / T, 2), end=" | ")
print(os.environ.get("IPP_B200_LIB", "default"), "ms/sim:", " ".join(f"{x:.4f}" for x in out), "checksum", chk)
