"""AC_IRL.train (forward solve with the reward net queried every step, ac_irl.py:634-732): episodes/s."""
import contextlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(0)
mat = rng.dirichlet(np.ones(15), size=21)
with contextlib.redirect_stdout(sys.stderr):
    ac = AC_IRL(theta=6.5, d=15, reg=sys.argv[1] if len(sys.argv) > 1 else "none", mat_pi0=mat, demonstrations=[], seed=1, net_seed=2)
    out = {}
    for fused in (True, False):
        ac.train(max_episodes=5, stop_criteria=-1, verbose=False, fused=fused)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        E = 2000 if fused else 200
        ac.train(max_episodes=E, stop_criteria=-1, verbose=False, fused=fused)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out["fused" if fused else "host_driven"] = {"episodes": E, "seconds": dt, "steps_per_s": E * 15 / dt, "us_per_step": 1e6 * dt / (E * 15)}
print(json.dumps(out))
