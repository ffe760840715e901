import contextlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(0)
mat = rng.dirichlet(np.ones(15), size=21)
with contextlib.redirect_stdout(sys.stderr):
    ac = AC_IRL(theta=6.5, d=15, reg="none", mat_pi0=mat, demonstrations=[], seed=1, net_seed=2)
    ac.list_demonstrations = ac.generate_trajectories(20)
    ac.list_generated = ac.generate_trajectories(50)
    for _ in range(5): ac.update_reward()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 200
    for _ in range(n): ac.update_reward()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("update_reward (5+5 trajectories, reference minibatch): %.1f us per update" % (1e6 * dt / n))
