"""The N > 1 path: sharding of independent populations + ONE all-reduce of the flat gradient buffer.

CPU part (gloo, world_size 2, runs anywhere): the host-side logic in discrete_mean_field_game_b200/parallel.py
-- shard ranges, the summed [2+F] buffer and the batch-mean update -- with the oracle standing in for the
per-shard kernel: two ranks on half the populations each must reproduce the single-process update bit for bit
up to summation order.

GPU part (marked gpu, needs >= 2 devices, NCCL): 2-rank train_batch / update_reward_batch equal the 1-GPU result
on the concatenated populations, because Philox streams are keyed by the GLOBAL population id.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_exactly():
    from discrete_mean_field_game_b200.parallel import shard_range
    for total in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from discrete_mean_field_game_b200 import parallel
from oracle import mfg_oracle as O
rank, world, _ = parallel.init_from_env("gloo")
assert parallel.world_info() == (rank, world)
d, T, B = 15, 5, 12
rng = np.random.RandomState(0)                       # every rank builds the same global problem
pi0 = rng.dirichlet(np.ones(d), size=B)
y = rng.gamma(2.0, size=(T, B, d, d))
w = rng.rand(O.num_features(d)); theta = 8.0
b, e = parallel.shard_range(B, rank, world)
ref = O.rollout_frozen(pi0[b:e], theta, 0.1, 1e4, y[:, b:e], w=w)          # this rank's shard
acc = torch.tensor(np.concatenate([[ref["G_theta"]], ref["G_w"], [ref["R"]]]))
parallel.allreduce_sum_(acc)
th, wn = parallel.ac_update_from_acc(torch.tensor(theta), torch.tensor(w), acc, 0.05, 0.01, B)
full = O.rollout_frozen(pi0, theta, 0.1, 1e4, y, w=w)                      # the single-process answer
np.testing.assert_allclose(acc[0].item(), full["G_theta"], rtol=1e-12)
np.testing.assert_allclose(acc[1:-1].numpy(), full["G_w"], rtol=1e-12, atol=1e-15)
np.testing.assert_allclose(acc[-1].item(), full["R"], rtol=1e-12)
np.testing.assert_allclose(th.item(), theta + 0.01 / B * full["G_theta"], rtol=1e-14)
np.testing.assert_allclose(wn.numpy(), w + 0.05 / B * full["G_w"], rtol=1e-14)
t = torch.tensor([float(rank)]); parallel.allreduce_max_(t); assert t.item() == world - 1
dist.barrier(); dist.destroy_process_group()
print("rank %%d ok" %% rank)
'''


def test_gloo_world2_sharded_update_equals_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from discrete_mean_field_game_b200 import parallel
from discrete_mean_field_game_b200.mfg_ac2 import actor_critic
from discrete_mean_field_game_b200.ac_irl import AC_IRL
rank, world, local = parallel.init_from_env("nccl")
dev = torch.device("cuda", local)
d, T, B = 15, 15, 4096
rng = np.random.RandomState(0)
pi0 = np.float32(rng.dirichlet(np.ones(d), size=B))
w0 = rng.rand(136, 1)
def run(group, lo, hi):
    ac = actor_critic(theta=8.0, shift=0.1, alpha_scale=1e4, d=d, mat_pi0=pi0[:4], device=dev, seed=77)
    ac.w = w0.copy()
    res = ac.train_batch(pi0[lo:hi], num_episodes=3, T=T, lr_critic=0.1, lr_actor=0.01, pop_offset=lo, group=group)
    return ac.theta, ac.w.ravel().copy(), res["mean_reward"]
b, e = parallel.shard_range(B, rank, world)
th2, w2, mr2 = run(None, b, e)                        # 2 ranks, half the populations each, all-reduce
th1, w1, mr1 = run(False, 0, B)                       # every rank alone on all populations
np.testing.assert_allclose(th2, th1, rtol=1e-12)
np.testing.assert_allclose(w2, w1, rtol=1e-11)
np.testing.assert_allclose(mr2, mr1, rtol=1e-10)
# uneven shards (B not a multiple of the world): the batch-mean scale comes from the all-reduced TOTAL
Bu = 4095
def run_u(group, lo, hi):
    ac = actor_critic(theta=8.0, shift=0.1, alpha_scale=1e4, d=d, mat_pi0=pi0[:4], device=dev, seed=78)
    ac.w = w0.copy()
    ac.train_batch(pi0[lo:hi], num_episodes=2, T=T, lr_critic=0.1, lr_actor=0.01, pop_offset=lo, group=group)
    return ac.theta, ac.w.ravel().copy()
bu, eu = parallel.shard_range(Bu, rank, world)
thu2, wu2 = run_u(None, bu, eu)
thu1, wu1 = run_u(False, 0, Bu)
np.testing.assert_allclose(thu2, thu1, rtol=1e-12)
np.testing.assert_allclose(wu2, wu1, rtol=1e-11)
# reward step: RANK-COUNT INVARIANT -- the ranks all-reduce raw sums (demonstration gradient for dL/dr = -1, the
# unnormalised generated gradient sum_j e^{R_j} dR_j, Z, sum r_demo, the counts) and apply 1/N_demo and 1/Z afterwards:
# the 2-rank step equals the 1-rank step on the concatenated batch (ac_irl.py:390-406 over ALL trajectories)
def irl(group, lo, hi):
    ac = AC_IRL(theta=6.5, d=d, reg="none", mat_pi0=pi0[:21], demonstrations=[], device=dev, seed=5, net_seed=2)
    ds, da = ac.generate_batch(64, theta=8.0)
    gs, ga = ac.generate_batch(64)
    loss = ac.update_reward_batch(ds[:15, lo:hi].reshape(-1, d).contiguous(), da[:, lo:hi].reshape(-1, d, d).contiguous(),
                                  gs[:15, lo:hi].reshape(-1, d).contiguous(), ga[:, lo:hi].reshape(-1, d, d).contiguous(),
                                  hi - lo, "time_major", group=group)
    return ac._last_grad.clone(), ac.reward_params.flat.clone(), loss.clone()
lo, hi = parallel.shard_range(63, rank, world)                           # 32 + 31 trajectories
g_dp, p_dp, l_dp = irl(None, lo, hi)
g_1, p_1, l_1 = irl(False, 0, 63)
assert (g_dp - g_1).abs().max().item() <= 2e-5 * g_1.abs().max().item() + 1e-8
assert ((p_dp - p_1).abs() <= 1e-6).float().mean().item() > 0.99 and (p_dp - p_1).abs().max().item() <= 2.1e-4
np.testing.assert_allclose(l_dp.cpu().numpy(), l_1.cpu().numpy(), rtol=1e-6)
gathered = [torch.empty_like(p_dp) for _ in range(world)]
dist.all_gather(gathered, p_dp)
assert torch.equal(gathered[0], gathered[1])                            # replicas stay identical
# the whole IRL training step (forward solve + reward update on its record, ONE all-reduce)
def irl_step(group, lo, hi):
    ac = AC_IRL(theta=6.5, d=d, reg="none", mat_pi0=pi0[:21], demonstrations=[], device=dev, seed=5, net_seed=2)
    ac.w = w0.copy()
    ds, da = ac.generate_batch(64, theta=8.0)
    res = ac.irl_step_batch(torch.as_tensor(pi0[lo:hi], device=dev), ds[:15, lo:hi].reshape(-1, d).contiguous(),
                            da[:, lo:hi].reshape(-1, d, d).contiguous(), hi - lo, episode=1, lr_critic=0.1,
                            lr_actor=0.01, pop_offset=lo, group=group)
    return ac.theta, ac.w.ravel().copy(), ac.reward_params.flat.clone(), res["loss"].clone()
t2, ww2, pp2, ll2 = irl_step(None, lo, hi)
t1, ww1, pp1, ll1 = irl_step(False, 0, 63)
np.testing.assert_allclose(t2, t1, rtol=1e-9)
np.testing.assert_allclose(ww2, ww1, rtol=1e-8)
assert ((pp2 - pp1).abs() <= 1e-6).float().mean().item() > 0.99 and (pp2 - pp1).abs().max().item() <= 2.1e-4
np.testing.assert_allclose(ll2.cpu().numpy(), ll1.cpu().numpy(), rtol=1e-6)
dist.barrier(); dist.destroy_process_group()
print("rank %%d ok" %% rank)
'''


@pytest.mark.gpu
def test_nccl_world2_train_batch_equals_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
