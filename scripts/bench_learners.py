"""Serial-learner throughput (config 1 semantics on the GPU): L independent learners x E episodes x 15 steps."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
d, T = 15, 15
rng = np.random.RandomState(0)
mat = torch.as_tensor(rng.dirichlet(np.ones(d), size=21), dtype=torch.float32, device=dev)
out = []
for L, E, layout in ((1, 2000, "groups"), (1, 2000, "cta"), (148, 500, "groups"), (148, 500, "cta"), (592, 200, "groups"),
                     (592, 200, "cta"), (1184, 100, "groups"), (1184, 100, "cta"), (2368, 100, "groups"), (2368, 100, "cta"), (4736, 50, "auto"), (65536, 20, "auto")):
    theta = torch.full((L,), 8.86349, dtype=torch.float64, device=dev)
    w = torch.rand((L, 136), dtype=torch.float64, device=dev)
    kw = dict(shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=1, layout=layout)
    engine.learners(theta, w, mat, 2, T, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    engine.learners(theta, w, mat, E, T, episode0=2, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append({"learners": L, "layout": layout, "episodes": E, "seconds": dt, "population_steps_per_s": L * E * T / dt,
                "steps_per_s_per_learner": E * T / dt})
    print(json.dumps(out[-1]), flush=True)
