import contextlib, sys, os, cProfile, pstats, io, numpy as np, torch
sys.path.insert(0, os.getcwd())
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rng = np.random.RandomState(5)
g = rng.standard_gamma(1.0, size=(21, 15)); mat = g / g.sum(1, keepdims=True)
with contextlib.redirect_stdout(sys.stderr):
    one = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="dropout_l1l2", n_fc3=8, n_fc4=4, mat_pi0=mat, demonstrations=[], device=torch.device("cuda:0"), seed=1, net_seed=2)
    one.list_demonstrations = one.generate_trajectories(20)
    one.list_generated = one.generate_trajectories(50)
    for _ in range(20): one.update_reward()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(500): one.update_reward()
    torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
import time
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(1000): one.update_reward()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("update_reward x 1000: host %.1f us per update, with the device drained %.1f us" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
