"""BASELINE config 5: data-parallel IRL training step on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_irl_dp.py [--log2-traj 16] [--iters 10]

Every rank owns 2^k generated + 2^k demonstration trajectories (x 15 transitions).  One iteration =
  (a) the batched forward solve of ac_irl.py over the rank's populations with the reward net in the loop
      (rollout + record -> r_net over all transitions -> TD sums -> (theta, w) update on the device);
  (b) its record is the rank's generated batch (no second rollout);
  (c) one reward update (reward-net backward over demo; ONE pass over the generated record that also hands back the
      rewards (a) needs; log-sum-exp loss over the LOCAL generated trajectories) and TF-Adam;
  with ONE all-reduce of the flat [2+F+|r_net|] buffer (AC_IRL.irl_step_batch).
Timed on the device (CUDA events), max over ranks; rank 0 prints one JSON line.
"""
import argparse
import contextlib
import json
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from discrete_mean_field_game_b200 import engine, parallel
from discrete_mean_field_game_b200.ac_irl import AC_IRL

ap = argparse.ArgumentParser()
ap.add_argument("--log2-traj", type=int, default=16)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D, T, M = 15, 15, 1 << a.log2_traj
F = D * (D + 1) // 2 + D + 1
rng = np.random.RandomState(5)
g = rng.standard_gamma(1.0, size=(64, D))
with contextlib.redirect_stdout(sys.stderr):
    irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="none", n_fc3=8, n_fc4=4,
                 mat_pi0=g / g.sum(1, keepdims=True), demonstrations=[], device=dev, seed=1 + rank, net_seed=2)
ds, da = irl.generate_batch(M, theta=8.06)                     # "demonstrations" of this rank
ds, da = ds[:T].reshape(-1, D), da.reshape(-1, D, D)
pi0 = ds[:M].contiguous()                                      # start states of the rank's populations


def iteration(k):
    # (a)+(b)+(c) in one call: the reward-net pass over the generated record serves the forward solve (r_gen) and
    # the reward update (gradient, loss); one all-reduce of the flat [2+F+|r_net|] buffer
    res = irl.irl_step_batch(pi0, ds, da, M, episode=1 + k, pop_offset=rank * M,
                             group=None if world == 1 else dist.group.WORLD)
    return res["loss"]


for k in range(2):
    iteration(k)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(a.iters):
    loss = iteration(2 + k)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
parallel.allreduce_max_(t)
ms = float(t[0]) / a.iters
if rank == 0:
    print(json.dumps({
        "config": "BASELINE configs[4]: data-parallel IRL training step (forward solve with the reward net in the loop, its "
                  "record = the generated batch, + reward update), %d GPUs x 2^%d trajectories x %d transitions"
                  % (world, a.log2_traj, T),
        "n_gpus": world, "trajectories_per_gpu": M, "ms_per_iteration": ms, "irl_iters_per_s": 1e3 / ms,
        "population_steps_per_s": 1.0 * M * T * world / (ms * 1e-3),
        "reward_update_transitions_per_s": 2.0 * M * T * world / (ms * 1e-3),
        "allreduces_per_iteration": 1 if world > 1 else 0, "loss": [float(x) for x in loss.cpu()],
        "theta": irl.theta,
    }))
if world > 1:
    dist.destroy_process_group()
