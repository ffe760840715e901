"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import contextlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from discrete_mean_field_game_b200 import engine
from discrete_mean_field_game_b200.ac_irl import AC_IRL

dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for d, B, T in ((15, 37, 5), (16, 33, 4)):
    F = d * (d + 1) // 2 + d + 1
    pi0 = torch.as_tensor(rng.dirichlet(np.ones(d), size=B), dtype=torch.float32, device=dev)
    w = torch.as_tensor(rng.rand(F), dtype=torch.float64, device=dev)
    engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=1, outputs=(), want_acc=True)                       # train
    engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=1, want_acc=True,
                   outputs=("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads", "pi_final"))  # record
    engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=1, outputs=("pi_final",))                 # rollout only
pi0 = torch.as_tensor(rng.dirichlet(np.ones(64), size=5), dtype=torch.float32, device=dev)
engine.rollout(pi0, 8.0, 0.1, 1e4, 3, seed=2, outputs=("states", "actions"))                                     # generic d
acts = engine.rollout(pi0[:, :15].contiguous(), 8.0, 0.1, 1e4, 3, seed=2, reward="none", outputs=("actions",))["actions"]
engine.synthetic_check(acts)
g = rng.standard_gamma(1.0, size=(8, 15))
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="dropout_l1l2", n_fc3=8, n_fc4=4,
                 mat_pi0=g / g.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1, net_seed=2)
ds, da = irl.generate_batch(24, theta=8.06)
gs, ga = irl.generate_batch(24)
irl.update_reward_batch(ds[:15].reshape(-1, 15), da.reshape(-1, 15, 15), gs[:15].reshape(-1, 15), ga.reshape(-1, 15, 15),
                        24, "time_major", group=False)
torch.cuda.synchronize()
print("sanitize_small: done")
