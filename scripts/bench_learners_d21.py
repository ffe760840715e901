import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
for d in (15, 21):
    rng = np.random.RandomState(0)
    mat = torch.as_tensor(rng.dirichlet(np.ones(d), size=21), dtype=torch.float32, device=dev)
    F = d * (d + 1) // 2 + d + 1
    for L, E, layout in ((1, 2000, "groups"), (1, 2000, "cta"), (444, 200, "groups"), (444, 200, "cta"), (8192, 50, "auto")):
        theta = torch.full((L,), 8.86349, dtype=torch.float64, device=dev)
        w = torch.rand((L, F), dtype=torch.float64, device=dev)
        kw = dict(shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=1, layout=layout)
        engine.learners(theta, w, mat, 2, 15, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.learners(theta, w, mat, E, 15, episode0=2, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("d=%d learners=%d layout=%s: %.3g population-steps/s (%.3g steps/s per learner)" % (
            d, L, layout, L * E * 15 / dt, E * 15 / dt))
