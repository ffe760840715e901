"""Driver for a shared-memory wavefront audit (ncu SourceCounters: actual vs ideal wavefronts per instruction) of the kernels
next to the headline one: a CTA-per-learner run, lane-group learners, the wide kernel + DMMA TD pass at d = 64, the 32-lane
v2 kernel at d = 21."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for d, L, layout in ((15, 1, "cta"), (15, 4096, "groups"), (21, 1, "cta")):
    F = d * (d + 1) // 2 + d + 1
    mat = torch.as_tensor(rng.dirichlet(np.ones(d), size=21), dtype=torch.float32, device=dev)
    th = torch.full((L,), 8.0, dtype=torch.float64, device=dev)
    ww = torch.rand((L, F), dtype=torch.float64, device=dev)
    engine.learners(th, ww, mat, 40, 15, shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.01, seed=1, layout=layout)
for d, B in ((64, 4096), (21, 8192)):
    F = d * (d + 1) // 2 + d + 1
    g = rng.standard_gamma(1.0, size=(B, d)).astype(np.float32)
    pi0 = torch.as_tensor(g / g.sum(1, keepdims=True), device=dev)
    w = torch.as_tensor(rng.rand(F), dtype=torch.float64, device=dev)
    engine.rollout(pi0, 8.86349, 0.16, 12000.0, 16, w=w, seed=5, outputs=(), want_acc=True)
torch.cuda.synchronize()
print("done")
