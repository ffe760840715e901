#!/usr/bin/env python
"""Headline benchmark: population-steps/s of the batched actor-critic train step at d = 15.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
2^20 independent populations per GPU x d = 15 topics x 16-step episodes, synthetic Dirichlet(1)
start states, in-kernel Philox noise.  One "step" of this benchmark = ONE EPISODE of the batched
trainer: every population runs 16 transitions of (sample P, pi' = P^T pi, closed-form reward,
TD error, policy-gradient and critic-gradient contributions) with frozen (theta, w), followed by
the reduction of the [2+F] gradient buffer, (N > 1: one NCCL all-reduce of it) and the update
of (theta, w) on the device.  population-steps per step = B * T per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-pops 20]

Prints ONE JSON line (rank 0).  `value` is timed on the device with inputs resident in HBM;
`e2e` goes through the public NumPy-in/NumPy-out API (`actor_critic.train_batch`) with the
start states copied from pinned host memory every step (double-buffered under the previous step's
kernel) and theta/w/mean reward of every step read back into pinned host memory.
`--impl reference` times the SAME workload on the box's host cores: the batched per-episode trainer restated in
NumPy (`oracle.mfg_oracle.train_batch_port`, kind "port" -- the reference is Python and its checkout does not exist
on the GPU box) on a bounded sample of the populations, one process per core; when the reference checkout IS present
(the build container) the line also carries the unmodified reference's own `mfg_ac2.actor_critic.train` throughput.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, T = 15, 16
THETA, SHIFT, ALPHA_SCALE = 8.86349, 0.16, 12000.0
LR_CRITIC, LR_ACTOR = 0.1, 0.1
METRIC, UNIT = "population-steps/sec (d=15)", "population-steps/s"
ALGO_BYTES_TRAIN = 16.0        # train, no recording: 2*4d/T + 8 B per population-step
ALGO_BYTES_RECORD = 4.0 * (D + D * D) + 4.0 * D / T   # rollout + record: 963.75 B per population-step


def bench_config(B, world):
    """The ONE workload both arms are quoted on (BASELINE.json configs[2] at its largest size)."""
    return {"workload": "configs[2]: batched actor-critic train step (a1-a7), %d populations/GPU x d=15 x "
                        "16-step episodes, per-episode batch-mean update" % B,
            "populations_per_gpu": B, "d": D, "T": T, "update": "per_episode",
            "noise": "philox4x32-7 (in-kernel Gamma sampler)",
            "l2": "flushed between timed iterations (256 MiB write)", "parallelism": "dp%d" % world}


def source_sha():
    """Hash of the sources the headline kernel is compiled from: profiles/kernel_constants.json records the hash its
    instruction count was measured at, so a stale constant is detected instead of silently reported."""
    import hashlib
    h = hashlib.sha256()
    for name in ("dmfg_math.cuh", "dmfg_rollout.cuh", "dmfg_rollout2.cuh"):
        with open(os.path.join(ROOT, "discrete_mean_field_game_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_profile_constants():
    """Instructions per population-step of the dominant kernel, from the committed ncu summary."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_constants.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled in-process
    every 20 ms (a sample costs ~50 us and is available from the first step on); `nvidia-smi -lms 100` -- which needs
    more than a short timed region to print its first line on an 8-GPU box -- is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag = None, None, [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [x.strip() for x in visible.split(",") if x.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            try:                                     # exact match whatever the visibility mask looks like
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(mx), pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = [x[0] for x in self.samples]
            reasons = set()
            for x in self.samples:
                for bit, name in self.NVML_REASONS.items():
                    if x[3] & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(x[1] for x in self.samples) if sm else None,
                    "power_w_max": max(x[2] for x in self.samples) if sm else None, "samples": len(sm),
                    "reasons": sorted(reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


def synthetic_pi0(B, seed):
    rng = np.random.RandomState(seed)
    g = rng.standard_gamma(1.0, size=(B, D)).astype(np.float64)
    return (g / g.sum(1, keepdims=True)).astype(np.float32)


# ----------------------------------------------------------------------------- CPU arm
CPU_SAMPLE_POPS = 1024          # populations per process per CPU "step" (a bounded sample of the 2^20)


def _cpu_worker_batch(args):
    """the GPU arm's workload on one core: batched per-episode actor-critic over a sample of populations"""
    seed, pops, episodes = args
    from oracle import mfg_oracle as O
    rng = np.random.RandomState(seed)
    g = rng.standard_gamma(1.0, size=(pops, D))
    pi0 = g / g.sum(1, keepdims=True)
    w0 = rng.rand(O.num_features(D))
    t0 = time.perf_counter()
    O.train_batch_port(pi0, THETA, w0, SHIFT, ALPHA_SCALE, episodes, T=T, lr_critic=LR_CRITIC, lr_actor=LR_ACTOR, rng=rng)
    return time.perf_counter() - t0


def _cpu_worker_serial(args):
    """the reference's own semantics (one population, per-step updates: mfg_ac2.py:448-539), oracle port"""
    seed, episodes = args
    from oracle import mfg_oracle as O
    np.random.seed(seed)
    mat = O.synthetic_start_states(n_rows=21, n_cols=20, d=D, seed=0)
    w0 = np.random.rand(O.num_features(D))
    t0 = time.perf_counter()
    O.train_serial(mat, THETA, w0, SHIFT, ALPHA_SCALE, episodes, lr_critic=LR_CRITIC, lr_actor=LR_ACTOR,
                   flavour="mfg_ac2", num_steps=T)
    return time.perf_counter() - t0


def _cpu_worker_reference(args):
    """the UNMODIFIED reference (mfg_ac2.actor_critic.train, 15 transitions per episode) -- build container only"""
    seed, episodes = args
    from oracle import ref_runner
    return ref_runner.time_train(episodes, d=D, theta=THETA, shift=SHIFT, alpha_scale=ALPHA_SCALE,
                                 lr_critic=LR_CRITIC, lr_actor=LR_ACTOR, seed=seed)[0]


def cpu_throughput(episodes, cores, pool, pops=CPU_SAMPLE_POPS):
    """population-steps/s of the batched port over all cores"""
    t0 = time.perf_counter()
    pool.map(_cpu_worker_batch, [(100 + i, pops, episodes) for i in range(cores)])
    dt = time.perf_counter() - t0
    return cores * pops * episodes * T / dt, dt


def cpu_serial_throughput(episodes, cores, pool):
    t0 = time.perf_counter()
    pool.map(_cpu_worker_serial, [(100 + i, episodes) for i in range(cores)])
    dt = time.perf_counter() - t0
    return cores * episodes * T / dt, dt


def cpu_reference_throughput(episodes, cores, pool):
    t0 = time.perf_counter()
    pool.map(_cpu_worker_reference, [(100 + i, episodes) for i in range(cores)])
    dt = time.perf_counter() - t0
    return cores * episodes * 15 / dt, dt


def reference_itself(cores, pool, seconds=4.0):
    """Throughput of the reference's own train() when its checkout is present, else None."""
    try:
        from oracle import ref_runner
        if not ref_runner.available():
            return None
        cpu_reference_throughput(5, cores, pool)
        episodes = int(os.environ.get("DMFG_BENCH_REF_EPISODES", max(20, int(seconds * 1800 / 15))))
        v, dt = cpu_reference_throughput(episodes, cores, pool)
        return {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "seconds": dt,
                "sample": "%d processes x %d episodes x 15 steps of the unmodified mfg_ac2.actor_critic.train "
                          "(per-step updates, one population per process), d=15" % (cores, episodes)}
    except Exception as exc:                                   # pragma: no cover
        return {"unavailable": repr(exc)}


def make_pool(cores):
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    return mp.get_context("fork").Pool(cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # one CPU "step" = `episodes` batched episodes over CPU_SAMPLE_POPS populations per process; sized from a
    # calibration run so that steps + warmup finish within ~2.5 minutes
    pool = make_pool(cores)
    v0, _ = cpu_throughput(1, cores, pool)                                 # workers import the oracle; calibration
    budget = min(150.0 / (args.steps + args.warmup), 20.0)
    episodes = max(1, int(budget * v0 / (cores * CPU_SAMPLE_POPS * T)))
    episodes = int(os.environ.get("DMFG_BENCH_CPU_EPISODES", episodes))    # tests shrink the sample
    for _ in range(args.warmup):
        cpu_throughput(1, cores, pool)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_throughput(episodes, cores, pool)
    dt = time.perf_counter() - t0
    value = args.steps * cores * CPU_SAMPLE_POPS * episodes * T / dt
    serial, _ = cpu_serial_throughput(max(2, int(os.environ.get("DMFG_BENCH_CPU_EPISODES", 60))), cores, pool)
    ref = reference_itself(cores, pool)
    pool.close()
    sample = ("%d processes x %d episodes x %d populations x %d steps of the batched per-episode trainer per bench "
              "step (oracle port train_batch_port, NumPy float64, np.random.gamma)" % (cores, episodes, CPU_SAMPLE_POPS, T))
    B = 1 << args.log2_pops
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(B, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "reference_semantics_port": {"value": serial, "unit": UNIT,
                                                      "what": "oracle port of mfg_ac2.train: ONE population per "
                                                              "process, per-step updates"},
                         "reference_itself": ref},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from discrete_mean_field_game_b200 import _lib, engine
    from discrete_mean_field_game_b200.mfg_ac2 import actor_critic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engine.require_cuda()

    B = 1 << args.log2_pops
    F = D * (D + 1) // 2 + D + 1
    pi0_host = torch.from_numpy(synthetic_pi0(B, seed=3 + rank)).pin_memory()
    pi0 = pi0_host.to(dev)
    rng = np.random.RandomState(0)
    w = torch.as_tensor(rng.rand(F), dtype=torch.float64, device=dev)
    theta = torch.tensor([THETA], dtype=torch.float64, device=dev)
    out = {"acc": torch.empty(2 + F, dtype=torch.float64, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    pop_offset = rank * B

    def step(episode, ev=None):
        lr_c = LR_CRITIC / (episode + 1.0)
        lr_a = LR_ACTOR / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
        if ev:
            ev[0].record()
        r = engine.rollout(pi0, 0.0, SHIFT, ALPHA_SCALE, T, w=w, theta_dev=theta, gamma=1.0, reward="ac2",
                           seed=1234, pop_offset=pop_offset, step_offset=episode * T, outputs=(),
                           want_acc=True, out=out)
        if ev:
            ev[1].record()
        if world > 1:
            dist.all_reduce(r["acc"])
        engine.apply_update(D, theta, w, r["acc"], lr_c, lr_a, 1.0 / (B * world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for e in range(args.warmup):
        step(e)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = int(_lib.load().dmfg_kernel_launches())
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                     # L2 flush between timed iterations (untimed)
        step_ev[k][0].record()
        step(args.warmup + k, kern_ev[k])
        step_ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    gpu_launches = int(_lib.load().dmfg_kernel_launches()) - launches0        # counted by the library's launch sites
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = sum(a.elapsed_time(b) for a, b in step_ev)
    ms_kern = sum(a.elapsed_time(b) for a, b in kern_ev) / args.steps
    t = torch.tensor([ms_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0]) / args.steps
    value = B * T * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API: host start states in, theta / w out, every step
    ac = actor_critic(theta=THETA, shift=SHIFT, alpha_scale=ALPHA_SCALE, d=D,
                      mat_pi0=np.full((1, D), 1.0 / D), device=dev, dtype="float32", seed=1234)
    ac.w = rng.rand(F, 1)
    # every step copies its start states from pinned host memory (double-buffered under the previous step's
    # kernel) and reads (theta, w, mean reward) of THAT step back into pinned host memory (history=True)
    e2e_steps = max(3, min(args.steps, 40))               # (the first copy of the double-buffered pipeline is not hidden)
    ac.train_batch(pi0_host, num_episodes=2, T=T, lr_critic=LR_CRITIC, lr_actor=LR_ACTOR, pop_offset=pop_offset,
                   history=True)
    barrier()
    t0 = time.perf_counter()
    res = ac.train_batch(pi0_host, num_episodes=e2e_steps, T=T, lr_critic=LR_CRITIC, lr_actor=LR_ACTOR,
                         pop_offset=pop_offset, first_episode=2, history=True)
    assert res["theta_history"].shape == (e2e_steps,) and np.isfinite(res["theta_history"]).all()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B * T * world * e2e_steps / float(te[0])

    # ---- multi-GPU correctness inside the bench run: the summed [2+F] buffer of a small batch computed on the shards
    # of all ranks (global population ids -> same Philox streams) against the same batch on ONE rank
    dp_check = None
    if world > 1:
        Bc = 4096
        pic = torch.as_tensor(synthetic_pi0(Bc, seed=11), device=dev)
        lo, hi = rank * Bc // world, (rank + 1) * Bc // world
        part = engine.rollout(pic[lo:hi].contiguous(), THETA, SHIFT, ALPHA_SCALE, T, w=w, seed=99, pop_offset=lo,
                              outputs=(), want_acc=True)["acc"].clone()
        dist.all_reduce(part)
        full = engine.rollout(pic, THETA, SHIFT, ALPHA_SCALE, T, w=w, seed=99, pop_offset=0, outputs=(),
                              want_acc=True)["acc"]
        err = ((part - full).abs() / (full.abs() + 1e-300)).max()
        err_abs = (part - full).abs().max()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        dp_check = {"what": "sum over ranks of the sharded [2+F] accumulators vs the unsharded batch on one rank",
                    "populations": Bc, "T": T, "max_rel_err": float(err), "max_abs_err": float(err_abs),
                    "ok": bool(float(err) < 1e-9)}

    # ---- the other two modes of config 3 (single GPU numbers, for the roofline discussion)
    # (single-GPU side numbers: measured in the N = 1 run only -- rank 0 alone would otherwise sit in them while the other
    # ranks wait at the barrier, and nothing in here may touch a collective)
    modes = {}
    if rank == 0 and world == 1:
        def timed(fn, n=3):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(n):
                flush.zero_()
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            return tot / n
        Br = min(B, 1 << 18)                                  # actions of 2^18 x 16 steps = 3.8 GB
        rec_out = {"states": torch.empty((T + 1, Br, D), device=dev), "actions": torch.empty((T, Br, D, D), device=dev)}
        ms_roll = timed(lambda: engine.rollout(pi0, THETA, SHIFT, ALPHA_SCALE, T, reward="none", seed=7,
                                               outputs=("pi_final",)))
        ms_rec = timed(lambda: engine.rollout(pi0[:Br], THETA, SHIFT, ALPHA_SCALE, T, reward="none", seed=7,
                                              outputs=("states", "actions"), out=rec_out))
        modes = {"rollout_only": {"value": B * T / (ms_roll * 1e-3), "unit": UNIT, "populations": B},
                 "rollout_record": {"value": Br * T / (ms_rec * 1e-3), "unit": UNIT, "populations": Br,
                                    "hbm_gbs": Br * T * ALGO_BYTES_RECORD / (ms_rec * 1e-3) / 1e9}}
        del rec_out
        # the reference's own run (BASELINE configs[0], what cpu_baseline / --impl reference time on the host): serial
        # learners with per-step online updates, mfg_ac2.train semantics -- ONE learner, and one per host core
        mat = torch.as_tensor(synthetic_pi0(21, seed=0), device=dev)
        serial = {}
        for L in (1, os.cpu_count() or 1):
            th = torch.full((L,), THETA, dtype=torch.float64, device=dev)
            ww = torch.rand((L, F), dtype=torch.float64, device=dev)
            kw = dict(shift=SHIFT, alpha_scale=ALPHA_SCALE, lr_critic=LR_CRITIC, lr_actor=LR_ACTOR, seed=3)
            E = 2000
            ms = timed(lambda: engine.learners(th, ww, mat, E, T, episode0=2, **kw))
            serial[L] = L * E * T / (ms * 1e-3)
        modes["serial_learners"] = {"value": serial[1], "unit": UNIT, "learners": 1,
                                    "one_per_host_core": {"learners": os.cpu_count() or 1,
                                                          "value": serial[os.cpu_count() or 1]},
                                    "kernel": "learner_cta_kernel<15,PHILOX>"}
        # IRL iterations/s (BASELINE config 2): 4096 demonstration + 4096 generated trajectories x 15 steps,
        # one update_reward-equivalent = r_net backward over demo (hands back r_demo) + one pass over the generated
        # batch (forward, loss weights, backward) + loss terms + Adam
        import contextlib
        from discrete_mean_field_game_b200.ac_irl import AC_IRL
        M = 4096
        with contextlib.redirect_stdout(sys.stderr):          # the class prints like the reference does
            irl = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="none", n_fc3=8, n_fc4=4,
                         mat_pi0=synthetic_pi0(64, seed=5), demonstrations=[], device=dev, seed=1, net_seed=2)
        ds, da = irl.generate_batch(M, theta=8.06)
        gs, ga = irl.generate_batch(M)
        ds, da = ds[:15].reshape(-1, D), da.reshape(-1, D, D)
        gs, ga = gs[:15].reshape(-1, D), ga.reshape(-1, D, D)
        upd = lambda: irl.update_reward_batch(ds, da, gs, ga, M, "time_major", group=False)
        for _ in range(3):
            upd()
        torch.cuda.synchronize()
        samples = []
        for _ in range(10):                                   # L2 flushed, one update per synchronisation: the median
            flush.zero_()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(); upd(); b_.record(); torch.cuda.synchronize()
            samples.append(a_.elapsed_time(b_))
        ms_irl = float(np.median(samples))
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(20):                                   # the outer loop's pattern: updates back to back
            upd()
        b_.record(); torch.cuda.synchronize()
        ms_b2b = a_.elapsed_time(b_) / 20
        modes["irl_update"] = {"value": 1e3 / ms_irl, "unit": "IRL iters/s", "demo_trajectories": M,
                               "generated_trajectories": M, "transitions_per_iter": 2 * M * 15,
                               "transitions_per_s": 2 * M * 15 / (ms_irl * 1e-3), "gpu_launches_per_iter": 3,
                               "us_samples_min_max": [1e3 * min(samples), 1e3 * max(samples)],
                               "back_to_back": {"value": 1e3 / ms_b2b, "us_per_iter": 1e3 * ms_b2b}}
        del ds, da, gs, ga
        # AC_IRL.train (ac_irl.py:634-732: ONE learner, the reward net queried at every transition, the reference's default
        # regulariser with dropout active) as one kernel -- dmfg_irl_learners -- and the reference's own 5 + 5 trajectory
        # reward update (ac_irl.py:804-846) through the class
        with contextlib.redirect_stdout(sys.stderr):
            one = AC_IRL(theta=8.64, shift=0, alpha_scale=1e4, d=D, reg="dropout_l1l2", n_fc3=8, n_fc4=4,
                         mat_pi0=synthetic_pi0(21, seed=5), demonstrations=[], device=dev, seed=1, net_seed=2)
            one.train(max_episodes=20, stop_criteria=-1, verbose=False)
            torch.cuda.synchronize()
            E = 2000
            dt = None
            for _ in range(3):             # one CTA on an idle GPU: the first run also pays the clock ramp -- best of 3
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                one.train(max_episodes=E, stop_criteria=-1, consecutive=E, verbose=False)
                b_.record(); torch.cuda.synchronize()
                dt = a_.elapsed_time(b_) * 1e-3 if dt is None else min(dt, a_.elapsed_time(b_) * 1e-3)
            modes["irl_learner"] = {"value": E * 15 / dt, "unit": UNIT, "learners": 1, "episodes": E,
                                    "kernel": "irl_learner_cta_kernel<15,PHILOX> (reward net + dropout in the loop)",
                                    "us_per_transition": 1e6 * dt / (E * 15)}
            one.list_demonstrations = one.generate_trajectories(20)
            one.list_generated = one.generate_trajectories(50)
            for _ in range(100):           # (the first draw of a trajectory uploads it into the device pool)
                one.update_reward()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(500):
                one.update_reward()
            torch.cuda.synchronize()
            modes["irl_update_minibatch"] = {"value": 500 / (time.perf_counter() - t0), "unit": "IRL iters/s",
                                             "demo_trajectories": 5, "generated_trajectories": 5,
                                             "what": "AC_IRL.update_reward as the reference calls it (host lists in, e2e)"}
        del one
        if not args.skip_big_modes:
            # config 1's semantics at scale: 2^16 INDEPENDENT learners, each the reference's serial loop (private theta, w,
            # per-step updates) -- the same algorithm the CPU reference runs, one population per process
            L = 1 << 16
            th = torch.full((L,), THETA, dtype=torch.float64, device=dev)
            ww = torch.rand((L, F), dtype=torch.float64, device=dev)
            E = 8
            ms = timed(lambda: engine.learners(th, ww, mat, E, T, episode0=2, shift=SHIFT, alpha_scale=ALPHA_SCALE,
                                               lr_critic=LR_CRITIC, lr_actor=LR_ACTOR, seed=3, want_total_reward=False))
            modes["independent_learners"] = {"value": L * E * T / (ms * 1e-3), "unit": UNIT, "learners": L,
                                             "episodes": E, "kernel": "learners_v2_kernel<15,16,PHILOX>",
                                             "semantics": "mfg_ac2.train per learner (per-step updates of w, theta)"}
            del th, ww
            # BASELINE config 4 as stated: d = 64 / 256, 2^20 populations on one GPU, full train step (rollout + TD sums)
            for dw in (64, 256):
                try:
                    Bw = 1 << 20
                    Fw = dw * (dw + 1) // 2 + dw + 1
                    rw = np.random.RandomState(dw)
                    gw = rw.standard_gamma(1.0, size=(1 << 14, dw)).astype(np.float32)
                    piw = torch.as_tensor(gw / gw.sum(1, keepdims=True), device=dev).repeat(Bw >> 14, 1).contiguous()
                    wd = torch.as_tensor(rw.rand(Fw), dtype=torch.float64, device=dev)
                    fn = lambda n=Bw: engine.rollout(piw[:n], THETA, SHIFT, ALPHA_SCALE, T, w=wd, seed=5, outputs=(),
                                                     want_acc=True)
                    fn(1 << 12); torch.cuda.synchronize()
                    if dw <= 64:
                        fn(); torch.cuda.synchronize()         # full-size warm-up: the first call allocates the 4.5 GB record
                                                               # (cudaMalloc inside the timed region made this number jump
                                                               # between 440 and 650 ms from run to run)
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a_.record(); fn(); b_.record(); torch.cuda.synchronize()
                    ms = a_.elapsed_time(b_)
                    modes["wide_d%d" % dw] = {"value": Bw * T / (ms * 1e-3), "unit": UNIT, "populations": Bw, "d": dw,
                                              "ms": ms, "sampled_elements_per_s": Bw * T * dw * dw / (ms * 1e-3),
                                              "what": "train step: rollout_wide_kernel + TD pass (DMMA), one launch chain"}
                    del piw, wd
                    engine._workspaces.clear()
                    torch.cuda.empty_cache()
                except Exception as exc:                       # pragma: no cover
                    modes["wide_d%d" % dw] = {"error": repr(exc)[:200]}
            # BASELINE config 5 per-GPU step at its stated size: one data-parallel IRL training step over 2^20 trajectories
            try:
                Bi = 1 << 20
                pii = torch.as_tensor(synthetic_pi0(1 << 14, seed=9), device=dev).repeat(Bi >> 14, 1).contiguous()
                Md = 4096
                dsd, dad = irl.generate_batch(Md, theta=8.06)
                dsd, dad = dsd[:15].reshape(-1, D).contiguous(), dad.reshape(-1, D, D)
                with contextlib.redirect_stdout(sys.stderr):
                    irl.irl_step_batch(pii[:1 << 12], dsd, dad, Md, episode=1)
                    torch.cuda.synchronize()
                    irl.irl_step_batch(pii, dsd, dad, Md, episode=1)      # full-size warm-up: allocates the 15 GB record once
                    torch.cuda.synchronize()
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a_.record(); irl.irl_step_batch(pii, dsd, dad, Md, episode=2); b_.record(); torch.cuda.synchronize()
                ms = a_.elapsed_time(b_)
                modes["irl_dp_step"] = {"value": Bi * 15 / (ms * 1e-3), "unit": UNIT, "trajectories_per_gpu": Bi,
                                        "ms": ms, "demo_trajectories": Md,
                                        "what": "AC_IRL.irl_step_batch: rollout+record, reward net fwd/bwd over the "
                                                "record, TD sums, (all-reduce), theta/w update, Adam"}
                del pii, dsd, dad
            except Exception as exc:                           # pragma: no cover
                modes["irl_dp_step"] = {"error": repr(exc)[:200]}
        del irl
        engine._workspaces.clear()
        torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    achieved = B * T * ALGO_BYTES_TRAIN / (ms_kern * 1e-3) / 1e9
    consts = load_profile_constants()
    sha = source_sha()
    fresh = consts.get("source_sha") == sha
    ipp = consts.get("train_inst_per_population_step") if fresh else None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    hbm = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
           "algorithmic_bytes_per_population_step": ALGO_BYTES_TRAIN,
           "traffic": consts.get("train_dram_bytes_per_launch") if fresh else None}
    # the binding roof of this kernel is instruction issue (4 warp instructions / clock / SM): the train step moves
    # 16 B per population-step (0.2 % of HBM) and executes ~800 warp instructions for it
    roofline = {"bound": "issue", "kernel": "rollout_v2_kernel<15,PHILOX,train>", "kernel_ms": ms_kern,
                "unit": "warp-inst/clk/SM", "peak": 4.0, "achieved": None, "frac": None,
                "traffic": hbm["traffic"], "hbm": hbm,
                "warp_inst_per_population_step": ipp, "source_sha": sha,
                "inst_count_source": consts.get("source") if fresh else
                "STALE: profiles/kernel_constants.json was measured at source_sha %s" % consts.get("source_sha")}
    if ipp and clocks and clocks.get("sm_mhz"):
        ipc = ipp * (B * T / (ms_kern * 1e-3)) / (sms * clocks["sm_mhz"] * 1e6)
        roofline["achieved"], roofline["frac"] = ipc, ipc / 4.0
    if "rollout_record" in modes:
        modes["rollout_record"]["hbm_frac"] = modes["rollout_record"]["hbm_gbs"] / peak
    for k in ("wide_d64", "wide_d256"):
        if k in modes and "ms" in modes[k]:
            dw = modes[k]["d"]
            modes[k]["hbm_gbs"] = modes[k]["populations"] * T * 4.0 * (dw + 2) / (modes[k]["ms"] * 1e-3) / 1e9
            modes[k]["hbm_frac"] = modes[k]["hbm_gbs"] / peak

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pool = make_pool(cores)
        v0, _ = cpu_throughput(1, cores, pool)                 # workers import the oracle; calibration
        episodes = int(os.environ.get("DMFG_BENCH_CPU_EPISODES",
                                      max(1, int(12.0 * v0 / (cores * CPU_SAMPLE_POPS * T)))))   # ~10-15 s
        v, dt = cpu_throughput(episodes, cores, pool)
        serial, _ = cpu_serial_throughput(max(2, int(os.environ.get("DMFG_BENCH_CPU_EPISODES", 60))), cores, pool)
        ref = reference_itself(cores, pool)
        pool.close()
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                        "sample": "%d processes x %d episodes x %d populations x %d steps of the batched per-episode "
                                  "trainer (oracle port train_batch_port: the GPU arm's workload on a sample of its "
                                  "populations), d=15" % (cores, episodes, CPU_SAMPLE_POPS, T),
                        "reference_semantics_port": {"value": serial, "unit": UNIT,
                                                     "what": "oracle port of mfg_ac2.train: ONE population per process, "
                                                             "per-step updates (compare modes.independent_learners)"},
                        "reference_itself": ref}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (transcendentals) + f64 (state, reductions)", "data": "synthetic",
        "config": bench_config(B, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 4,
                "d2h_bytes_per_step": (F + 1) * 8 + 8, "steps": e2e_steps},
        "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "dp_check": dp_check, "modes": modes, "wall_s": t_wall,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    # ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version there when NCCL_DEBUG is
    # set on the box) are sent to stderr for the whole run; the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-pops", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-big-modes", action="store_true", help="skip the 2^20-population side modes (configs 4, 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
