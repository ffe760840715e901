import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
d, T = 15, 16
rng = np.random.RandomState(0)
mat = torch.as_tensor(rng.dirichlet(np.ones(d), size=21), dtype=torch.float32, device=dev)
L, E = 65536, 8
theta = torch.full((L,), 8.86349, dtype=torch.float64, device=dev)
w = torch.rand((L, 136), dtype=torch.float64, device=dev)
kw = dict(shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=1, layout="groups")
engine.learners(theta, w, mat, 2, T, **kw)
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); engine.learners(theta, w, mat, E, T, episode0=2, **kw); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(os.environ.get("DMFG_LIB_PATH", "default"), "independent learners: %.3f ms, %.3e population-steps/s" % (best, L * E * T / (best * 1e-3)))
