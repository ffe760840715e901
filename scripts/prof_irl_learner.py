import contextlib, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(0)
mat = rng.dirichlet(np.ones(15), size=21)
with contextlib.redirect_stdout(sys.stderr):
    ac = AC_IRL(theta=6.5, d=15, reg=sys.argv[1] if len(sys.argv) > 1 else "dropout_l1l2", mat_pi0=mat, demonstrations=[], seed=1, net_seed=2)
    ac.train(max_episodes=5, stop_criteria=-1, verbose=False, fused=True)
    ac.train(max_episodes=400, stop_criteria=-1, consecutive=400, verbose=False, fused=True)
torch.cuda.synchronize()
print("done")
