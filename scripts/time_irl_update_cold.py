"""One reward update (BASELINE config 2) timed the way bench.py's irl_update mode does it -- L2 flushed, one update per
synchronisation -- for the finishing-launch form and the six-launch chain, next to the back-to-back figure."""
import contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200.ac_irl import AC_IRL
dev = torch.device("cuda:0")
M, D = 4096, 15
rng = np.random.RandomState(5)
g = rng.standard_gamma(1.0, size=(64, D)); mat = g / g.sum(1, keepdims=True)
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="none", n_fc3=8, n_fc4=4, mat_pi0=mat, demonstrations=[],
                 device=dev, seed=1, net_seed=2)
ds, da = irl.generate_batch(M, theta=8.06)
gs, ga = irl.generate_batch(M)
ds, da = ds[:15].reshape(-1, D).contiguous(), da.reshape(-1, D, D)
gs, ga = gs[:15].reshape(-1, D).contiguous(), ga.reshape(-1, D, D)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False)

def cold(n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))

def warm(n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n

for mode in (True, "chain", True, "chain"):
    irl.fused_reward_step = mode
    c = cold(); w = warm()
    print("fused_reward_step=%-5s  flushed, one per sync: median %.1f us (min %.1f) | back to back: %.1f us (%.0f it/s)" % (mode, c[0], c[1], w, 1e6 / w))
