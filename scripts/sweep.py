"""BASELINE configs 3 and 4 on one GPU: population-steps/s of the three modes over B and d.

    python scripts/sweep.py > gpurun_out/sweep.json
config 3: B in 2^16..2^20, d = 15, T = 16 (rollout only / rollout + record / full train step)
config 4: d in {64, 256}, generic kernel (warp per population); B bounded so a run stays in seconds
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from discrete_mean_field_game_b200 import engine

dev = torch.device("cuda:0")
T = 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


def pi0_of(B, d, seed=3):
    g = torch.empty((B, d), device=dev).exponential_(generator=torch.Generator(device=dev).manual_seed(seed))
    return (g / g.sum(1, keepdim=True)).contiguous()


rows = []
for d, Bs in ((15, [1 << k for k in range(16, 21)]), (16, [1 << 18]), (64, [1 << 12, 1 << 14, 1 << 16]), (256, [1 << 8, 1 << 10, 1 << 12])):
    F = d * (d + 1) // 2 + d + 1
    w = torch.rand(F, dtype=torch.float64, device=dev)
    for B in Bs:
        pi0 = pi0_of(B, d)
        r = {"d": d, "B": B, "T": T}
        ms = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("pi_final",)))
        r["rollout_only"] = B * T / (ms * 1e-3)
        ms = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=7, outputs=(), want_acc=True))
        r["train_step"] = B * T / (ms * 1e-3)
        rec_bytes = 4.0 * (d + d * d) * T * B
        if rec_bytes < 40e9:
            out = {"states": torch.empty((T + 1, B, d), device=dev), "actions": torch.empty((T, B, d, d), device=dev)}
            ms = timed(lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7,
                                              outputs=("states", "actions"), out=out))
            r["rollout_record"] = B * T / (ms * 1e-3)
            r["record_GBps"] = rec_bytes / (ms * 1e-3) / 1e9
            del out
        rows.append(r)
        print(json.dumps(r), file=sys.stderr, flush=True)
print(json.dumps(rows))
