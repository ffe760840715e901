import os, sys, time
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bench.py") else os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for d, B in ((64, 65536), (256, 4096)):
    T = 16
    F = d * (d + 1) // 2 + d + 1
    g = torch.rand((B, d), device=dev) + 0.01
    pi0 = (g / g.sum(1, keepdim=True)).contiguous()
    w = torch.rand(F, dtype=torch.float64, device=dev)
    rec = engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="ac2", seed=7, outputs=("states", "rewards", "grads"))
    ms_roll_nograd = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("pi_final",)))
    ms_roll = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="ac2", seed=7, outputs=("states", "rewards", "grads")))
    ms_td = timed(lambda: engine.td_accumulate(rec["states"], rec["rewards"], rec["grads"], w, want_deltas=False))
    ms_full = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=7, outputs=(), want_acc=True))
    print("d=%d B=%d: rollout(no grad) %.2f ms, rollout(grad, states) %.2f ms, td_accumulate %.2f ms, train step %.2f ms" % (d, B, ms_roll_nograd, ms_roll, ms_td, ms_full))
