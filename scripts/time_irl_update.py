"""Times one reward update (BASELINE config 2: 4096 + 4096 trajectories x 15 steps) and its two reward-net launches
with CUDA events; A/B tool (DMFG_LIB_PATH selects the library)."""
import contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from discrete_mean_field_game_b200 import engine
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
p = irl.reward_params
dconst = torch.full((ds.shape[0],), -1.0 / M, device=dev)

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3

t_upd = timed(lambda: irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False))
irl.fused_reward_step = "chain"
t_chain = timed(lambda: irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False))
irl.fused_reward_step = True
t_bwd = timed(lambda: engine.rnet_backward(p.flat, ds, da, dconst, 8, 4))
t_fwd = timed(lambda: engine.rnet_forward(p.flat, ds, da, 8, 4))
r_demo = engine.rnet_forward(p.flat, ds, da, 8, 4)
t_gen = timed(lambda: engine.rnet_backward_gen(p.flat, gs, ga, 8, 4, 15, r_demo, M))
print("%s: update %.1f us (%.0f it/s; six-launch chain %.1f us) | backward %.1f us | backward_gen %.1f us | forward %.1f us | %d transitions per launch" % (
    os.environ.get("DMFG_LIB_PATH", "default"), t_upd, 1e6 / t_upd, t_chain, t_bwd, t_gen, t_fwd, ds.shape[0]))
