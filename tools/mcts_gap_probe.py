"""How far are the GPU search's root visit counts from the fp64 oracle's on the 'deeper problem' of tests/test_gpu_mcts.py?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mcts_oracle as morc
from tests._util import engine_cfg, make_params, oracle_cfg
from ipp_rl_b200 import BatchedEngine
from ipp_rl_b200.planning.mcts_zero import BatchedMCTS

X = Y = 24
params = make_params(X, Y, 1.0, 8, 20, 6)
cfg = oracle_cfg(params)
for seed, T in ((21, 12), (22, 24), (23, 24)):
    hyper = dict(puct_init=6.0, puct_base=10000, num_mcts_simulations=48, gamma=0.95, dirichlet_alpha=0.3, dirichlet_eps=0.0,
                 forced_playout_factor=2.0, max_valid_action_distance=7.5)
    H = 3
    meta = dict(episode_horizon=H, scenario_info=None)
    rng = np.random.RandomState(seed)
    var0 = rng.uniform(0.1, 2.0, (T, Y, X)).astype(np.float32)
    prev = np.stack([rng.randint(2, X - 2, T) + 0.5, rng.randint(2, Y - 2, T) + 0.5, rng.choice([8.0, 14.0, 20.0], T)], axis=1)
    budgets = rng.uniform(6.0, 40.0, T).astype(np.float32)
    num_actions = 3 * X * Y

    def policy_of(prev_pose, budget):
        k = int(round(prev_pose[0] * 3 + prev_pose[1] * 5 + prev_pose[2])) + int(budget * 4)
        a = np.arange(num_actions, dtype=np.float64)
        s = np.sin(a * 0.731 + k * 1.37) * 1000.0
        pol = 0.05 + (s - np.floor(s)) ** 4
        return (pol / pol.sum()).astype(np.float32), np.float32(0.02 * (k % 11))

    with BatchedEngine(engine_cfg(params, T, layout=2)) as eng:
        eng.reset(0.5, 1.0)
        eng.set_state(var=var0)
        with BatchedMCTS(eng, hyper, meta) as mcts:
            def ev(leaf):
                pri = np.zeros((T, num_actions), np.float32)
                val = np.zeros(T, np.float32)
                for t in range(T):
                    if leaf.kind[t] != 1:
                        continue
                    pp = prev[t] if leaf.level[t] < 0 else np.array([leaf.col[t] + 0.5, leaf.row[t] + 0.5, eng.altitudes[leaf.level[t]]])
                    pri[t], val[t] = policy_of(pp, float(leaf.budget[t]))
                return pri, val
            mcts.begin(budgets, prev)
            for _ in range(48):
                mcts.simulate(ev)
            st = mcts.root_stats()
    gaps = []
    for t in range(T):
        o = morc.OracleMCTS(cfg, hyper, H, evaluator=lambda info: policy_of(info["previous_action"], np.float32(info["budget"])))
        o.search(var0[t].astype(np.float64), prev[t], float(budgets[t]), 48)
        dense_n = np.zeros(num_actions)
        ok = st["action_ids"][t] >= 0
        dense_n[st["action_ids"][t][ok]] = st["Nsa"][t][ok]
        gaps.append(int(np.abs(dense_n - o.Nsa[()]).sum()))
    print("seed", seed, "T", T, "sum|dN| per tree:", gaps, "of", int(st["Ns"][0]), "visits")
