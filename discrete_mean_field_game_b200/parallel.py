"""Data parallelism over independent populations: host-side plumbing only.

Populations (episodes) are independent given the shared parameters (theta, w, reward net), so every
rank owns a contiguous shard and the only exchange is ONE sum of a small flat buffer per update
([2+F] doubles for the actor-critic step, [|r_net|] floats for the reward step) -- SURVEY 8(e),
DESIGN.md section 6.  Philox streams are keyed by the GLOBAL population id (``pop_offset``), so the
trajectories -- and therefore the summed gradients -- do not depend on the number of ranks.

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the box; gloo in the CPU tests, where
the same functions run on CPU tensors).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info(group=None):
    """(rank, world) of the default / given process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and group is not False:
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(total, rank, world):
    """Contiguous block [begin, end) of `total` populations owned by `rank`; sizes differ by at most one,
    the first `total % world` ranks take the extra element."""
    total, rank, world = int(total), int(rank), int(world)
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_sum_(buf, group=None):
    """In-place sum of `buf` over the ranks (no-op for a single process).  The gradient buffers are tiny
    (<= 16 KB): the call is latency-bound, one per update."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def total_count(n, group=None, device=None):
    """Sum of the per-rank counts `n` as a Python int (one tiny blocking all-reduce, done ONCE per train call):
    shard_range hands out shards whose sizes differ by one when total % world != 0, so the batch-mean scale must
    come from the TOTAL, not from local_count * world."""
    _, world = world_info(group)
    if world == 1:
        return int(n)
    backend = dist.get_backend(group)
    dev = device if (device is not None and backend == "nccl") else torch.device("cpu")
    t = torch.tensor([int(n)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0])


def allreduce_max_(buf, group=None):
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
    return buf


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_*); returns (rank, world, local_rank).  Single process when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def ac_update_from_acc(theta, w, acc, lr_critic_eff, lr_actor_eff, total_populations):
    """Host-tensor form of dmfg_ac_apply_update (mfg_ac2.py:511-522) for the batch-mean update:
    theta += lr_a/B * acc[0]; w += lr_c/B * acc[1:1+F].  Used by the CPU tests of the reduction
    semantics; the product applies the same update on the device."""
    F = w.numel()
    theta = theta + lr_actor_eff / total_populations * acc[0]
    w = w + lr_critic_eff / total_populations * acc[1:1 + F]
    return theta, w
