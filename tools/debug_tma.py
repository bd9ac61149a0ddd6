import sys, numpy as np
sys.path.insert(0, '.')
from ipp_rl_b200 import BatchedEngine, EngineConfig
B=int(sys.argv[1]) if len(sys.argv)>1 else 8
X=int(sys.argv[2]) if len(sys.argv)>2 else 24
cfg = EngineConfig(batch=B, x_dim=X, y_dim=X, resolution=1.0, min_altitude=8.0, max_altitude=20.0, altitude_spacing=6.0, layout=1, seed=7)
eng = BatchedEngine(cfg)
print('tma_active', eng.tma_active)
eng.reset(0.5, 1.82); eng.synth_ground_truth(1)
ids = np.random.RandomState(0).randint(0, eng.num_actions, B).astype(np.int32)
r = eng.step(ids)
print('rewards', r[:8])
