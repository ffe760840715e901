"""Times the headline kernel (train step / record / rollout only) with CUDA events: A/B tool for kernel builds
(DMFG_LIB_PATH selects the library)."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from discrete_mean_field_game_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--log2-pops", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--d", type=int, default=15)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, T, D = 1 << a.log2_pops, 16, a.d
F = D * (D + 1) // 2 + D + 1
rng = np.random.RandomState(3)
g = rng.standard_gamma(1.0, size=(B, D))
pi0 = torch.as_tensor(g / g.sum(1, keepdims=True), dtype=torch.float32, device=dev)
w = torch.as_tensor(rng.rand(F), dtype=torch.float64, device=dev)
Br = min(B, 1 << 18)
rec_out = {"states": torch.empty((T + 1, Br, D), device=dev), "actions": torch.empty((T, Br, D, D), device=dev)}
modes = {
    "train": lambda i: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, w=w, seed=1234, step_offset=i * T, outputs=(), want_acc=True),
    "rollout": lambda i: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("pi_final",)),
    "record": lambda i: engine.rollout(pi0[:Br], 8.86349, 0.16, 12000.0, T, reward="none", seed=7, outputs=("states", "actions"), out=rec_out),
}
res = {}
for name, fn in modes.items():
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for i in range(a.iters):
        ev[i][0].record(); fn(i); ev[i][1].record()
    torch.cuda.synchronize()
    ms = float(np.median([x.elapsed_time(y) for x, y in ev]))
    n = (Br if name == "record" else B) * T
    res[name] = (ms, n / ms / 1e3)
chk = engine.rollout(pi0[:4096], 8.86349, 0.16, 12000.0, T, w=w, seed=1234, outputs=(), want_acc=True)["acc"].cpu().numpy()
import hashlib
print(os.environ.get("DMFG_LIB_PATH", "default"), "acc-sha", hashlib.sha1(chk.tobytes()).hexdigest()[:10], " ".join("%s %.3f ms %.1f Mps/s" % (k, v[0], v[1] / 1e3) for k, v in res.items()))
