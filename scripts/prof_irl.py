"""Tiny driver for ncu: a few IRL reward updates at BASELINE config 2 (4096 + 4096 trajectories x 15 steps)."""
import contextlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL

M, D = 4096, 15
dev = torch.device("cuda:0")
rng = np.random.RandomState(5)
g = rng.standard_gamma(1.0, size=(64, D))
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="none", n_fc3=8, n_fc4=4,
                 mat_pi0=g / g.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1, net_seed=2)
irl.one_pass_reward_update = os.environ.get("DMFG_IRL_ONE_PASS", "1") != "0"      # A/B: the three-launch chain
ds, da = irl.generate_batch(M, theta=8.06)
gs, ga = irl.generate_batch(M)
ds, da = ds[:15].reshape(-1, D), da.reshape(-1, D, D)
gs, ga = gs[:15].reshape(-1, D), ga.reshape(-1, D, D)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False)
e1.record()
torch.cuda.synchronize()
print("done: %.1f us per reward update (%d + %d trajectories)" % (1e3 * e0.elapsed_time(e1) / n, M, M))
