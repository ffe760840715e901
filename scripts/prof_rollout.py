"""Tiny driver for ncu: a few launches of the headline kernel (train step, no recording)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from discrete_mean_field_game_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--log2-pops", type=int, default=20)
ap.add_argument("--mode", default="train", choices=["train", "record", "rollout"])
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--d", type=int, default=15)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, T, D = 1 << a.log2_pops, 16, a.d
rng = np.random.RandomState(3)
g = rng.standard_gamma(1.0, size=(B, D))
pi0 = torch.as_tensor(g / g.sum(1, keepdims=True), dtype=torch.float32, device=dev)
w = torch.as_tensor(rng.rand(D * (D + 1) // 2 + D + 1), dtype=torch.float64, device=dev)
for i in range(a.iters):
    if a.mode == "train":
        engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=1234, step_offset=i * T, outputs=(), want_acc=True)
    elif a.mode == "record":
        engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("states", "actions"))
    else:
        engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("pi_final",))
torch.cuda.synchronize()
print("done")
