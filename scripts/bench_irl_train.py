"""AC_IRL.train (forward solve with the reward net queried every step, ac_irl.py:634-732): steps/s of the fused kernel
(dmfg_irl_learners; best of 5 runs of 2000 episodes, CUDA-event timed) and of the host-driven chain."""
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
        E = 2000 if fused else 200
        best = None
        for _ in range(5 if fused else 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            ac.train(max_episodes=E, stop_criteria=-1, consecutive=E, verbose=False, fused=fused)
            b.record(); torch.cuda.synchronize()
            dt = a.elapsed_time(b) * 1e-3
            best = dt if best is None else min(best, dt)
        out["fused" if fused else "host_driven"] = {"episodes": E, "seconds": best, "steps_per_s": E * 15 / best, "us_per_step": 1e6 * best / (E * 15)}
print(json.dumps(out))
