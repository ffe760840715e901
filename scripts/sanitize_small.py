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
# independent serial learners (v2 math at d = 15 / 16, first-generation kernel at d = 4 and in float64)
for d, dt in ((15, torch.float32), (16, torch.float32), (21, torch.float32), (4, torch.float32), (15, torch.float64)):
    F = d * (d + 1) // 2 + d + 1
    L = 19
    mat = torch.as_tensor(rng.dirichlet(np.ones(d), size=7), dtype=dt, device=dev)
    th = torch.full((L,), 8.0, dtype=torch.float64, device=dev)
    ww = torch.rand((L, F), dtype=torch.float64, device=dev)
    engine.learners(th, ww, mat, 3, 5, shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=1,
                    layout="groups")
    if dt == torch.float32 and d in (15, 16, 21):                                                           # CTA per learner
        engine.learners(th, ww, mat, 3, 5, shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=1,
                        layout="cta", trace=True)
pi0 = torch.as_tensor(rng.dirichlet(np.ones(64), size=5), dtype=torch.float32, device=dev)
engine.rollout(pi0, 8.0, 0.1, 1e4, 3, seed=2, outputs=("states", "actions"))                                     # wide kernel, d = 64
for dd, var in ((21, "auto"), (32, "auto"), (21, "generic"), (47, "auto"), (130, "auto"), (32, "generic"), (64, "auto")):
    # auto at d = 21 / 32: the 32-lane v2 kernel; generic there and the other d: the wide kernel (1 / 1 / 4 pairs per
    # lane), with the TD pass on the DMMA kernels at d = 32 / 64
    pw = torch.as_tensor(rng.dirichlet(np.ones(dd), size=6), dtype=torch.float32, device=dev)
    ww = torch.as_tensor(rng.rand(dd * (dd + 1) // 2 + dd + 1), dtype=torch.float64, device=dev)
    engine.rollout(pw, 8.0, 0.1, 1e4, 3, w=ww, seed=2, want_acc=True, variant=var,
                   outputs=("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads", "pi_final"))
    engine.rollout(pw, 8.0, 0.1, 1e4, 3, w=ww, seed=2, want_acc=True, variant=var, outputs=())
acts = engine.rollout(pi0[:, :15].contiguous(), 8.0, 0.1, 1e4, 3, seed=2, reward="none", outputs=("actions",))["actions"]
engine.synthetic_check(acts)
g = rng.standard_gamma(1.0, size=(8, 15))
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="dropout_l1l2", n_fc3=8, n_fc4=4,
                 mat_pi0=g / g.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1, net_seed=2)
ds, da = irl.generate_batch(24, theta=8.06)
gs, ga = irl.generate_batch(24)
for one_pass in (True, False):            # generated half in one launch (rnet_kernel<TRAJ>) / forward -> loss -> backward
    irl.one_pass_reward_update = one_pass
    irl.update_reward_batch(ds[:15].reshape(-1, 15), da.reshape(-1, 15, 15), gs[:15].reshape(-1, 15),
                            ga.reshape(-1, 15, 15), 24, "time_major", group=False)
# ---- round 2: the fused per-step launch (dmfg_ac_step), AC_IRL.train as one kernel (dmfg_irl_learners), the reward net at
# d = 16 (generic instantiation) and d = 21 (32 lanes per transition), TMEM-parked accumulators, tensor-core fc3 gradient
from discrete_mean_field_game_b200 import mfg_ac2
gm = rng.standard_gamma(1.0, size=(21, 15))
mat = gm / gm.sum(1, keepdims=True)
with contextlib.redirect_stdout(sys.stderr):
    ac = mfg_ac2.actor_critic(theta=8.0, shift=0.16, alpha_scale=12000, d=15, mat_pi0=mat, dtype="float32", seed=21)
    ac.train_batch(np.float32(rng.dirichlet(np.ones(15), size=77)), num_episodes=1, T=3, lr_critic=0.1, lr_actor=0.01,
                   update="per_step")
    irl.train(max_episodes=3, stop_criteria=-1, verbose=False, use_graph=False)
    # the record-based TD pass at d = 15 (td_delta_small_kernel + td_gw_kernel) through one data-parallel IRL step
    irl_nd = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=15, reg="none", n_fc3=8, n_fc4=4,
                    mat_pi0=g / g.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1, net_seed=2)
    irl_nd.irl_step_batch(np.float32(rng.dirichlet(np.ones(15), size=133)), ds[:15].reshape(-1, 15), da.reshape(-1, 15, 15), 24)
    for dd in (16, 21):
        gg = rng.standard_gamma(1.0, size=(8, dd))
        irl2 = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=dd, reg="none", n_fc3=6, n_fc4=3,
                      mat_pi0=gg / gg.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1, net_seed=2)
        s2, a2 = irl2.generate_batch(20)
        irl2.update_reward_batch(s2[:15].reshape(-1, dd), a2.reshape(-1, dd, dd), s2[:15].reshape(-1, dd),
                                 a2.reshape(-1, dd, dd), 20, "time_major", group=False)
    # update_reward on trajectories resident in the device pool (gathered batches: rnet_kernel<..., GATHER> in both backward
    # forms + irl_step_finish_kernel), with pool growth, and the gathered forward
    irl.one_pass_reward_update = True
    irl.list_demonstrations = irl.generate_trajectories(7)
    irl.list_generated = irl.generate_trajectories(70)
    for _ in range(16):
        irl.update_reward()
    pool = irl._pool
    from discrete_mean_field_game_b200 import engine
    engine.rnet_forward(irl.reward_params.flat, pool["states"], pool["actions"], 8, 4, gather=(15, [3, 0, 5]))
torch.cuda.synchronize()
print("sanitize_small: done")
