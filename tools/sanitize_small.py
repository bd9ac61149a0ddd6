"""Small, fast exercise of every hand-written kernel for compute-sanitizer (memcheck / racecheck / synccheck runs are 10-100x
slower than native, so the sizes are tiny but cover: all five layouts, the three step kernels (bulk / cp.async / LSU), both
reward modes, the adaptive mask, host noise + measurement read-back, clipped / corner footprints, predict (persistent and
job-list), path rollouts, the lock-step MCTS kernels (memoised rollouts), eval, observe, GRF / field / prior reset, the host step
with pinned buffers (ids fetched by the persistent kernel, rewards written in place).

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ipp_rl_b200 import BatchedEngine, EngineConfig  # noqa: E402
from ipp_rl_b200.engine import pinned_array  # noqa: E402
from ipp_rl_b200.planning.mcts_zero import BatchedMCTS  # noqa: E402


def main():
    X, Y, B = 48, 40, 600  # > 148 SMs x 4 warps: the ticket counter and the plan rings wrap a few times
    rng = np.random.RandomState(0)
    gt = rng.uniform(0, 1, (B, Y, X)).astype(np.float32)
    layouts = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4]
    ids_pin, out_pin = pinned_array((B,), np.int32), pinned_array((B,), np.float32)
    for layout in layouts:
        cfg = EngineConfig(batch=B, x_dim=X, y_dim=Y, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0,
                           layout=layout, seed=7, interval_factor=0.25, value_threshold=0.45)
        with BatchedEngine(cfg) as eng:
            for path in ("async", "lsu"):
                eng.set_step_path(path)
                eng.reset(0.5, 1.82)
                eng.set_ground_truth(gt)
                for t in range(4):
                    ids = rng.randint(0, eng.num_actions, B).astype(np.int32)
                    ids[:6] = [0, X - 1, X * (Y - 1), X * Y - 1, eng.num_actions - 1, X * Y]
                    if t == 2:
                        noise = rng.standard_normal((B, eng.max_measurements)).astype(np.float32)
                        eng.step(ids, noise=noise, reward_mode=t & 1, adaptive=True, return_measurements=True)
                    else:
                        eng.step(ids, reward_mode=t & 1, adaptive=(t == 3))
                for rep in range(2):  # pinned host buffers: ids fetched by the kernel (bulk path), rewards in place
                    ids_pin[:] = rng.randint(0, eng.num_actions, B)
                    eng.step(ids_pin, out=out_pin)
                eng.predict(rng.randint(0, eng.num_actions, B).astype(np.int32), commit=True)
                eng.predict(rng.randint(0, eng.num_actions, B).astype(np.int32), commit=True, adaptive=True)
                eng.predict(rng.randint(0, eng.num_actions, B).astype(np.int32), commit=False, reward_mode=1)
                eng.predict(rng.randint(0, eng.num_actions, 50).astype(np.int32), env_index=rng.randint(0, B, 50).astype(np.int32),
                            prev_poses=np.array([3.0, 4.0, 9.0]), commit=False)
                eng.rollout(rng.randint(0, eng.num_actions, (40, 4)).astype(np.int32), env_index=rng.randint(0, B, 40).astype(np.int32))
            eng.eval()
            eng.observe(0, 4)
            eng.generate_ground_truth(3.0, seed=1, first_env=0, n_env=8)
            eng.generate_field("hotspot_random_field", 5, seed=3)
            eng.generate_field("split_random_field", 5, seed=4, first_env=3, n_env=50)
            eng.reset_shuffled(0.5, fit_gaussian_process=True, scale=1.82, seed=2)
            eng.reset_shuffled(0.5, fit_gaussian_process=False, scale=0.5, seed=2)
        # the tree search needs a square grid (the reference's action ids are a bijection only there)
        cfg = EngineConfig(batch=96, x_dim=40, y_dim=40, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0,
                           layout=layout, seed=7, interval_factor=0.25, value_threshold=0.45)
        with BatchedEngine(cfg) as eng:
            eng.reset(0.5, 1.82)
            eng.set_ground_truth(gt[:96, :40, :40].copy())
            eng.step(rng.randint(0, eng.num_actions, 96).astype(np.int32))
            hyper = dict(puct_init=6.0, puct_base=10000, num_mcts_simulations=6, gamma=0.95, dirichlet_alpha=0.3, dirichlet_eps=0.25,
                         forced_playout_factor=2.0, max_valid_action_distance=7.5)
            with BatchedMCTS(eng, hyper, dict(episode_horizon=3, scenario_info=None), n_trees=64) as mcts:
                mcts.get_policy(np.full(64, 30.0, np.float32), None, evaluator=None, temperature=1, deploy_time=False,
                                rng=np.random.default_rng(1))
            print(f"layout {layout}: ok, launches {eng.launches}", flush=True)


if __name__ == "__main__":
    main()
