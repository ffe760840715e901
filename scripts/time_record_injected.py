import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
dev = torch.device("cuda:0")
B, T, D = 1 << 18, 16, 15
rng = np.random.RandomState(3)
g = rng.standard_gamma(1.0, size=(B, D)).astype(np.float32)
pi0 = torch.as_tensor(g / g.sum(1, keepdims=True), device=dev)
y = torch.empty((T, B, D, D), dtype=torch.float32, device=dev).exponential_() + 0.1      # positive variates standing in for Gamma draws
out = {"states": torch.empty((T + 1, B, D), device=dev), "actions": torch.empty((T, B, D, D), device=dev)}
fn = lambda: engine.rollout(pi0, 8.86349, 0.16, 12000.0, T, reward="none", noise_y=y, outputs=("states", "actions"), out=out)
for _ in range(2): fn()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
bytes_ = B * T * (2 * D * D * 4 + D * 4)
print("injected-noise record: %.3f ms, %.3e population-steps/s, %.0f GB/s (read y + write P, states)" % (best, B * T / (best * 1e-3), bytes_ / (best * 1e-3) / 1e9))
