"""Thin device-side plumbing over libdmfg: torch owns device memory and streams,
the C ABI does the work.  Every function here takes/returns CUDA tensors in the
library's time-major layouts (states [T+1,B,d], actions [T,B,d,d], per-step
scalars [T,B]); nothing falls back to the host.
"""
from __future__ import annotations

import collections
import ctypes as C

import torch

from . import _lib
from ._lib import (DISCOUNT_CUMULATIVE, DISCOUNT_STEP, F32, F64, NOISE_ACTIONS, NOISE_INJECTED, NOISE_PHILOX,
                   REWARD_KINDS, VARIANTS, LearnersArgs, RolloutArgs, TdArgs, check, num_features)

ROLLOUT_OUTPUTS = ("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads", "pi_final")

_workspaces = collections.OrderedDict()      # (device index, stream handle) -> scratch tensor, least recently used first
_MAX_WORKSPACES = 16


def _dtype_code(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError("stream dtype must be torch.float32 or torch.float64, got %r" % (dtype,))


def _raw_stream(device):
    """cudaStream_t of torch's current stream on `device` as an int (the raw query: torch.cuda.current_stream() builds a
    Stream object per call, ~5 us -- a fifth of a small reward update's host time)"""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class _on:
    """`with _on(device):` == `with torch.cuda.device(device)`, free when the device is already current (the context
    manager costs ~6 us per call, four times per small reward update)"""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def _stream_ptr(device):
    return C.c_void_p(_raw_stream(device))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require(t, name, device, dtype, shape):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor" % name)
    if t.device != device:
        raise ValueError("%s is on %s, expected %s" % (name, t.device, device))
    if t.dtype != dtype:
        raise TypeError("%s has dtype %s, expected %s" % (name, t.dtype, dtype))
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _workspace(device, nbytes):
    """Grow-only scratch per (device, stream)."""
    if nbytes == 0:
        return None
    key = (device.index, _raw_stream(device))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    _workspaces.move_to_end(key)
    while len(_workspaces) > _MAX_WORKSPACES:            # short-lived side streams must not pin scratch forever
        _workspaces.popitem(last=False)
    return ws


def drop_workspace(device, stream):
    """Forget the scratch cached for `stream` (a side / capture stream that is going away)."""
    _workspaces.pop((torch.device(device).index, stream.cuda_stream), None)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("discrete_mean_field_game_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    _lib.load()


def rollout(pi0, theta, shift, alpha_scale, T, *, w=None, gamma=1.0, reward="ac2", discount="step",
            noise_y=None, seed=0, pop_offset=0, step_offset=0, outputs=("states",), want_acc=False,
            rewards_in=None, theta_dev=None, variant="auto", out=None, actions_in=None, step_offset_dev=None):
    """dmfg_rollout on CUDA tensors.

    pi0 [B,d] (float32/float64 selects the stream dtype); noise_y [T,B,d,d] selects
    injected noise, otherwise in-kernel Philox keyed by (seed, pop_offset + b).
    Returns a dict with the requested `outputs` (time-major tensors) and, with
    want_acc, 'acc' = [sum delta*g, sum delta*phi (F), sum r] as float64.
    `out` may carry preallocated tensors for any output (reused across calls).
    actions_in [1,B,d,d] (with T == 1) evaluates GIVEN transition matrices instead of sampling.
    """
    lib = _lib.load()
    if not pi0.is_cuda:
        raise TypeError("pi0 must be a CUDA tensor")
    device, dtype = pi0.device, pi0.dtype
    B, d = pi0.shape
    F = num_features(d)
    pi0 = _require(pi0, "pi0", device, dtype, (B, d))
    a = RolloutArgs()
    a.struct_size = C.sizeof(RolloutArgs)
    a.dtype, a.d, a.T, a.B, a.pop_offset = _dtype_code(dtype), d, int(T), B, int(pop_offset)
    a.theta, a.shift, a.alpha_scale, a.gamma = float(theta), float(shift), float(alpha_scale), float(gamma)
    a.theta_dev = _ptr(_require(theta_dev, "theta_dev", device, torch.float64, (1,))) if theta_dev is not None else None
    a.reward_kind = REWARD_KINDS[reward]
    a.discount_kind = DISCOUNT_STEP if discount == "step" else DISCOUNT_CUMULATIVE
    a.variant = VARIANTS[variant]
    a.seed, a.step_offset = int(seed) & (2 ** 64 - 1), int(step_offset)
    if step_offset_dev is not None:          # device scalar added to step_offset (CUDA-graph replays)
        a.step_offset_dev = _ptr(_require(step_offset_dev, "step_offset_dev", device, torch.int64, (1,)))
    if actions_in is not None:
        a.noise_kind = NOISE_ACTIONS
        a.noise_y = _ptr(_require(actions_in, "actions_in", device, dtype, (T, B, d, d)))
    elif noise_y is not None:
        a.noise_kind = NOISE_INJECTED
        a.noise_y = _ptr(_require(noise_y, "noise_y", device, dtype, (T, B, d, d)))
    else:
        a.noise_kind = NOISE_PHILOX
    a.pi0 = _ptr(pi0)
    if w is not None:
        a.w = _ptr(_require(w, "w", device, torch.float64, (F,)))
    if rewards_in is not None:
        a.rewards_in = _ptr(_require(rewards_in, "rewards_in", device, dtype, (T, B)))
    shapes = dict(states=(T + 1, B, d), actions=(T, B, d, d), alpha=(T, B, d, d), alpha_deriv=(T, B, d, d),
                  rewards=(T, B), deltas=(T, B), grads=(T, B), pi_final=(B, d))
    res = {}
    with _on(device):
        for name in outputs:
            if name not in shapes:
                raise ValueError("unknown output %r" % name)
            t = out.get(name) if out else None
            if t is None:
                t = torch.empty(shapes[name], dtype=dtype, device=device)
            else:
                _require(t, name, device, dtype, shapes[name])
            res[name] = t
            setattr(a, name, _ptr(t))
        if want_acc:
            acc = out.get("acc") if out else None
            if acc is None:
                acc = torch.empty(2 + F, dtype=torch.float64, device=device)
            res["acc"] = _require(acc, "acc", device, torch.float64, (2 + F,))
            a.acc = _ptr(acc)
        need = lib.dmfg_rollout_workspace_bytes(C.byref(a))
        ws = _workspace(device, need)
        if ws is not None:
            a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_rollout(C.byref(a), _stream_ptr(device)))
    return res


AC_STEP_D = (15, 16, 21)


def ac_step(pi, theta_dev, w, shift, alpha_scale, lr_critic_eff, lr_actor_eff, scale, *, gamma=1.0, reward="ac2",
            noise_y=None, seed=0, pop_offset=0, step_offset=0, lr_dev=None, out=None, step_offset_dev=None):
    """dmfg_ac_step: ONE transition of B populations + the batch-mean update of (theta_dev, w) in a single launch
    (float32 streams, d in AC_STEP_D).  Returns dict(pi_final [B,d], acc [2+F])."""
    lib = _lib.load()
    device = pi.device
    B, d = pi.shape
    F = num_features(d)
    a = RolloutArgs()
    a.struct_size = C.sizeof(RolloutArgs)
    a.dtype, a.d, a.T, a.B, a.pop_offset = _dtype_code(pi.dtype), d, 1, B, int(pop_offset)
    a.shift, a.alpha_scale, a.gamma = float(shift), float(alpha_scale), float(gamma)
    a.reward_kind, a.discount_kind = REWARD_KINDS[reward], DISCOUNT_STEP
    a.seed, a.step_offset = int(seed) & (2 ** 64 - 1), int(step_offset)
    if step_offset_dev is not None:
        a.step_offset_dev = _ptr(_require(step_offset_dev, "step_offset_dev", device, torch.int64, (1,)))
    if noise_y is not None:
        a.noise_kind = NOISE_INJECTED
        a.noise_y = _ptr(_require(noise_y, "noise_y", device, pi.dtype, (1, B, d, d)))
    else:
        a.noise_kind = NOISE_PHILOX
    a.pi0 = _ptr(_require(pi, "pi", device, pi.dtype, (B, d)))
    _require(theta_dev, "theta_dev", device, torch.float64, (1,))
    _require(w, "w", device, torch.float64, (F,))
    a.w = _ptr(w)
    res = {}
    with _on(device):
        res["pi_final"] = (out or {}).get("pi_final")
        if res["pi_final"] is None:
            res["pi_final"] = torch.empty((B, d), dtype=pi.dtype, device=device)
        res["acc"] = (out or {}).get("acc")
        if res["acc"] is None:
            res["acc"] = torch.empty(2 + F, dtype=torch.float64, device=device)
        a.pi_final, a.acc = _ptr(res["pi_final"]), _ptr(res["acc"])
        ws = _workspace(device, lib.dmfg_ac_step_workspace_bytes(C.byref(a)))
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_ac_step(C.byref(a), _ptr(theta_dev), _ptr(w), float(lr_critic_eff), float(lr_actor_eff),
                               float(scale), _ptr(lr_dev) if lr_dev is not None else None, _stream_ptr(device)))
    return res


def td_accumulate(states, rewards, grads, w, *, gamma=1.0, discount="step", want_deltas=True, want_acc=True):
    """dmfg_td_accumulate: TD errors / accumulators from a recorded batch of trajectories."""
    lib = _lib.load()
    device, dtype = states.device, states.dtype
    T1, B, d = states.shape
    T = T1 - 1
    F = num_features(d)
    a = TdArgs()
    a.struct_size = C.sizeof(TdArgs)
    a.dtype, a.d, a.T, a.B = _dtype_code(dtype), d, T, B
    a.gamma = float(gamma)
    a.discount_kind = DISCOUNT_STEP if discount == "step" else DISCOUNT_CUMULATIVE
    a.states = _ptr(_require(states, "states", device, dtype, (T + 1, B, d)))
    a.rewards = _ptr(_require(rewards, "rewards", device, dtype, (T, B)))
    if grads is not None:
        a.grads = _ptr(_require(grads, "grads", device, dtype, (T, B)))
    a.w = _ptr(_require(w, "w", device, torch.float64, (F,)))
    res = {}
    with _on(device):
        if want_deltas:
            res["deltas"] = torch.empty((T, B), dtype=dtype, device=device)
            a.deltas = _ptr(res["deltas"])
        if want_acc:
            res["acc"] = torch.empty(2 + F, dtype=torch.float64, device=device)
            a.acc = _ptr(res["acc"])
        ws = _workspace(device, lib.dmfg_td_workspace_bytes(C.byref(a)))
        if ws is not None:
            a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_td_accumulate(C.byref(a), _stream_ptr(device)))
    return res


def critic_eval(states, w=None, want_features=True, want_values=False):
    """calc_features / calc_value (mfg_ac2.py:290-344) for states [N,d] on device."""
    lib = _lib.load()
    device, dtype = states.device, states.dtype
    N, d = states.shape
    F = num_features(d)
    states = _require(states, "states", device, dtype, (N, d))
    res = {}
    with _on(device):
        if want_features:
            res["features"] = torch.empty((N, F), dtype=dtype, device=device)
        if want_values:
            res["values"] = torch.empty((N,), dtype=dtype, device=device)
            w = _require(w, "w", device, torch.float64, (F,))
        check(lib.dmfg_critic_eval(_dtype_code(dtype), d, N, _ptr(states), _ptr(w) if want_values else None,
                                   _ptr(res.get("features")), _ptr(res.get("values")), _stream_ptr(device)))
    return res


def traj_metrics(generated, empirical, time_major=True):
    """L1 and Jensen-Shannon divergence per (trajectory, hour) (mfg_ac2.py:546-563, 627-650).
    generated: rollout states [H,B,d] (time_major) or [B,H,d]; empirical [B,H,d].  Returns (l1, jsd) [B,H] float64."""
    lib = _lib.load()
    device, dtype = generated.device, generated.dtype
    if time_major:
        H, B, d = generated.shape
        gsb, gsh = d, B * d
    else:
        B, H, d = generated.shape
        gsb, gsh = H * d, d
    _require(generated, "generated", device, dtype, tuple(generated.shape))
    _require(empirical, "empirical", device, dtype, (B, H, d))
    with _on(device):
        l1 = torch.empty((B, H), dtype=torch.float64, device=device)
        js = torch.empty((B, H), dtype=torch.float64, device=device)
        check(lib.dmfg_traj_metrics(_dtype_code(dtype), d, B, H, _ptr(generated), gsb, gsh, _ptr(empirical), H * d, d,
                                    _ptr(l1), _ptr(js), _stream_ptr(device)))
    return l1, js


def synthetic_check(actions, want_jsd=True):
    """Analytic check of recorded trajectories against the MFG backward equation (mfg_synthetic.py:726-899).
    actions: time-major rollout record [T,B,d,d].  Returns (l1, jsd) [B,T] float64 (jsd None if not wanted)."""
    lib = _lib.load()
    device, dtype = actions.device, actions.dtype
    T, B, d, d2 = actions.shape
    if d != d2:
        raise ValueError("actions must be [T,B,d,d]")
    _require(actions, "actions", device, dtype, (T, B, d, d))
    with _on(device):
        l1 = torch.empty((B, T), dtype=torch.float64, device=device)
        js = torch.empty((B, T), dtype=torch.float64, device=device) if want_jsd else None
        check(lib.dmfg_synthetic_check(_dtype_code(dtype), d, B, T, _ptr(actions), _ptr(l1),
                                       _ptr(js) if want_jsd else None, _stream_ptr(device)))
    return l1, js


def apply_update(d, theta_dev, w, acc, lr_critic_eff, lr_actor_eff, scale, lr_dev=None):
    """theta += lr_a*scale*acc[0]; w += lr_c*scale*acc[1:1+F]  (mfg_ac2.py:511-522), on device.
    ``lr_dev`` [2] float64 device tensor (lr_critic_eff, lr_actor_eff) replaces the two by-value step sizes."""
    lib = _lib.load()
    device = w.device
    with _on(device):
        if lr_dev is not None:
            check(lib.dmfg_ac_apply_update_dev(int(d), _ptr(theta_dev), _ptr(w), _ptr(acc),
                                               _ptr(_require(lr_dev, "lr_dev", device, torch.float64, (2,))),
                                               float(scale), _stream_ptr(device)))
            return
        check(lib.dmfg_ac_apply_update(int(d), _ptr(theta_dev), _ptr(w), _ptr(acc), float(lr_critic_eff),
                                       float(lr_actor_eff), float(scale), _stream_ptr(device)))


def learners(theta, w, mat_pi0, E, T, *, shift, alpha_scale, episode0=0, gamma=1.0, lr_critic=0.1,
             lr_actor=0.001, constant=False, reward="ac2", discount="step", start_rows=None, noise_y=None,
             seed=0, learner_offset=0, noise_episode_offset=0, trace=False, want_total_reward=True, layout="auto"):
    """dmfg_ac_learners: L independent serial learners with per-step updates.

    theta [L] float64 and w [L,F] float64 are updated IN PLACE.  shift / alpha_scale may be
    floats or [L] float64 tensors.  mat_pi0 [S,d] selects the stream dtype.  layout: "auto" | "groups" (16 / 32
    lanes per learner, throughput form) | "cta" (a CTA per learner: latency form, float32, d = 15 / 16 / 21).
    """
    lib = _lib.load()
    device, dtype = mat_pi0.device, mat_pi0.dtype
    S, d = mat_pi0.shape
    L = theta.shape[0]
    F = num_features(d)
    a = LearnersArgs()
    a.struct_size = C.sizeof(LearnersArgs)
    a.dtype, a.d, a.T, a.L, a.E = _dtype_code(dtype), d, int(T), L, int(E)
    a.learner_offset, a.episode0 = int(learner_offset), int(episode0)
    a.theta = _ptr(_require(theta, "theta", device, torch.float64, (L,)))
    a.w = _ptr(_require(w, "w", device, torch.float64, (L, F)))
    if isinstance(shift, torch.Tensor):
        a.shift = _ptr(_require(shift, "shift", device, torch.float64, (L,)))
    else:
        a.shift_scalar = float(shift)
    if isinstance(alpha_scale, torch.Tensor):
        a.alpha_scale = _ptr(_require(alpha_scale, "alpha_scale", device, torch.float64, (L,)))
    else:
        a.alpha_scale_scalar = float(alpha_scale)
    a.gamma, a.lr_critic, a.lr_actor = float(gamma), float(lr_critic), float(lr_actor)
    a.constant_lr = 1 if constant else 0
    a.reward_kind = REWARD_KINDS[reward]
    a.discount_kind = DISCOUNT_STEP if discount == "step" else DISCOUNT_CUMULATIVE
    a.mat_pi0, a.S = _ptr(_require(mat_pi0, "mat_pi0", device, dtype, (S, d))), S
    a.seed = int(seed) & (2 ** 64 - 1)
    a.noise_episode_offset = int(noise_episode_offset)
    a.layout = {"auto": 0, "groups": 1, "cta": 2}[layout]
    if noise_y is not None:
        a.noise_kind = NOISE_INJECTED
        a.noise_y = _ptr(_require(noise_y, "noise_y", device, dtype, (L, E, T, d, d)))
        if start_rows is None:
            raise ValueError("injected noise needs start_rows [L,E]")
    else:
        a.noise_kind = NOISE_PHILOX
    if start_rows is not None:
        a.start_rows = _ptr(_require(start_rows, "start_rows", device, torch.int32, (L, E)))
    res = {}
    with _on(device):
        if trace:
            res["theta_trace"] = torch.empty((L, E, T), dtype=torch.float64, device=device)
            res["delta_trace"] = torch.empty((L, E, T), dtype=torch.float64, device=device)
            a.theta_trace, a.delta_trace = _ptr(res["theta_trace"]), _ptr(res["delta_trace"])
        if want_total_reward:
            res["total_reward"] = torch.empty((L, E), dtype=torch.float64, device=device)
            a.total_reward = _ptr(res["total_reward"])
        res["pi_final"] = torch.empty((L, d), dtype=dtype, device=device)
        a.pi_final = _ptr(res["pi_final"])
        check(lib.dmfg_ac_learners(C.byref(a), _stream_ptr(device)))
    return res


def irl_learners(theta, w, mat_pi0, E, T, params, n_fc3, n_fc4, *, shift, alpha_scale, episode0=1, gamma=1.0,
                 lr_critic=0.1, lr_actor=0.001, constant=False, discount="cumulative", start_rows=None, noise_y=None,
                 seed=0, learner_offset=0, noise_episode_offset=0, dropout_seed=None, keep_prob=0.4, sample_offset=0,
                 trace=False):
    """dmfg_irl_learners: AC_IRL.train (ac_irl.py:634-732) as ONE kernel -- L serial learners (a CTA each) whose reward
    is r_net(pi_t, P_t) evaluated inside the loop.  theta [L] / w [L,F] float64 are updated in place; `params` is the flat
    float32 reward net; float32 streams, d in {15, 16}.  dropout_seed keys the in-kernel Philox dropout masks
    (None: no dropout); transition (l, e, t) uses sample id sample_offset + (l*E + e)*T + t."""
    from ._lib import DROPOUT_NONE, DROPOUT_PHILOX, IrlNetArgs
    lib = _lib.load()
    device, dtype = mat_pi0.device, mat_pi0.dtype
    S, d = mat_pi0.shape
    L = theta.shape[0]
    F = num_features(d)
    a = LearnersArgs()
    a.struct_size = C.sizeof(LearnersArgs)
    a.dtype, a.d, a.T, a.L, a.E = _dtype_code(dtype), d, int(T), L, int(E)
    a.learner_offset, a.episode0 = int(learner_offset), int(episode0)
    a.theta = _ptr(_require(theta, "theta", device, torch.float64, (L,)))
    a.w = _ptr(_require(w, "w", device, torch.float64, (L, F)))
    a.shift_scalar, a.alpha_scale_scalar = float(shift), float(alpha_scale)
    a.gamma, a.lr_critic, a.lr_actor = float(gamma), float(lr_critic), float(lr_actor)
    a.constant_lr = 1 if constant else 0
    a.reward_kind = REWARD_KINDS["none"]
    a.discount_kind = DISCOUNT_STEP if discount == "step" else DISCOUNT_CUMULATIVE
    a.mat_pi0, a.S = _ptr(_require(mat_pi0, "mat_pi0", device, dtype, (S, d))), S
    a.seed = int(seed) & (2 ** 64 - 1)
    a.noise_episode_offset = int(noise_episode_offset)
    if noise_y is not None:
        a.noise_kind = NOISE_INJECTED
        a.noise_y = _ptr(_require(noise_y, "noise_y", device, dtype, (L, E, T, d, d)))
        if start_rows is None:
            raise ValueError("injected noise needs start_rows [L,E]")
    else:
        a.noise_kind = NOISE_PHILOX
    if start_rows is not None:
        a.start_rows = _ptr(_require(start_rows, "start_rows", device, torch.int32, (L, E)))
    n = IrlNetArgs()
    n.struct_size = C.sizeof(IrlNetArgs)
    n.n_fc3, n.n_fc4 = int(n_fc3), int(n_fc4)
    n.params = _ptr(_require(params, "params", device, torch.float32, (rnet_param_count(d, n_fc3, n_fc4),)))
    n.keep_prob = float(keep_prob)
    if dropout_seed is None:
        n.dropout = DROPOUT_NONE
    else:
        n.dropout, n.seed, n.sample_offset = DROPOUT_PHILOX, int(dropout_seed) & (2 ** 64 - 1), int(sample_offset)
    res = {}
    with _on(device):
        if trace:
            res["theta_trace"] = torch.empty((L, E, T), dtype=torch.float64, device=device)
            res["delta_trace"] = torch.empty((L, E, T), dtype=torch.float64, device=device)
            res["reward_trace"] = torch.empty((L, E, T), dtype=torch.float32, device=device)
            a.theta_trace, a.delta_trace = _ptr(res["theta_trace"]), _ptr(res["delta_trace"])
            n.reward_trace = _ptr(res["reward_trace"])
        res["total_reward"] = torch.empty((L, E), dtype=torch.float64, device=device)
        a.total_reward = _ptr(res["total_reward"])
        res["pi_final"] = torch.empty((L, d), dtype=dtype, device=device)
        a.pi_final = _ptr(res["pi_final"])
        check(lib.dmfg_irl_learners(C.byref(a), C.byref(n), _stream_ptr(device)))
    return res


def gamma_sample(shape, seed=0, pop=0):
    """Gamma(shape,1) variates from the kernels' own sampler (testing aid)."""
    lib = _lib.load()
    shape = shape.contiguous().to(torch.float32)
    out = torch.empty_like(shape)
    with torch.cuda.device(shape.device):
        check(lib.dmfg_gamma_sample(_ptr(shape), shape.numel(), int(seed), int(pop), _ptr(out),
                                    _stream_ptr(shape.device)))
    return out


def digamma(x):
    """Device digamma of the library for float32/float64 tensors (testing aid)."""
    lib = _lib.load()
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.dmfg_digamma(_dtype_code(x.dtype), _ptr(x), x.numel(), _ptr(out), _stream_ptr(x.device)))
    return out


def philox(ctr, key, gamma_stream=False):
    """Host evaluation of the library's Philox4x32-10 (known-answer tests); gamma_stream=True evaluates the
    block function of the Gamma sampler (dmfg_gamma_philox_rounds() rounds) instead."""
    lib = _lib.load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    (lib.dmfg_philox4x32_gamma if gamma_stream else lib.dmfg_philox4x32_10)(c, k, o)
    return tuple(int(v) for v in o)


def umma_probe(A, B, a_cfg, b_cfg, a_desc=None, b_desc=None):
    """ONE kind::tf32 MMA: out[128,16] = A[128,8] . B[8,16]; *_cfg = (mn_major, lbo_bytes, sbo_bytes) place the operands,
    *_desc = (lbo, sbo) are written into the descriptors (default: the same values) (testing aid)."""
    a_desc = a_desc or a_cfg[1:]
    b_desc = b_desc or b_cfg[1:]
    lib = _lib.load()
    device = A.device
    _require(A, "A", device, torch.float32, (128, 8))
    _require(B, "B", device, torch.float32, (8, 16))
    with _on(device):
        out = torch.empty((128, 32), dtype=torch.float32, device=device)
        check(lib.dmfg_umma_probe(_ptr(A), _ptr(B), int(a_cfg[0]), int(a_cfg[1]), int(a_cfg[2]), int(b_cfg[0]),
                                  int(b_cfg[1]), int(b_cfg[2]), _ptr(out), _stream_ptr(device), int(a_desc[0]),
                                  int(a_desc[1]), int(b_desc[0]), int(b_desc[1])))
    return out


def umma_selftest(h, z):
    """out[512, 8] = sum_p h[p]^T z[p] on the tcgen05 / TMEM path of the fc3 weight gradient (testing aid).
    h [P,16,512], z [P,16,8] float32 CUDA tensors."""
    lib = _lib.load()
    device = h.device
    P = h.shape[0]
    _require(h, "h", device, torch.float32, (P, 16, 512))
    _require(z, "z", device, torch.float32, (P, 16, 8))
    with _on(device):
        out = torch.empty((512, 8), dtype=torch.float32, device=device)
        check(lib.dmfg_umma_selftest(_ptr(h), _ptr(z), P, _ptr(out), _stream_ptr(device)))
    return out


def gamma_philox_rounds():
    return int(_lib.load().dmfg_gamma_philox_rounds())


# ----------------------------------------------------------------------------- IRL path (a10-a13)
_param_counts = {}


def rnet_param_count(d, n_fc3, n_fc4):
    key = (int(d), int(n_fc3), int(n_fc4))
    n = _param_counts.get(key)
    if n is None:
        n = _param_counts[key] = int(_lib.load().dmfg_rnet_param_count(*key))
    return n


def rnet_param_offsets(d, n_fc3, n_fc4):
    """Offsets of conv1/w, conv1/b, conv2/w, conv2/b, fc3/w, fc3/b, fc4/w, fc4/b, out/w, out/b."""
    o = (C.c_int64 * 10)()
    check(_lib.load().dmfg_rnet_param_offsets(int(d), int(n_fc3), int(n_fc4), o))
    return [int(x) for x in o]


def _rnet_args(params, states, actions, n_fc3, n_fc4, mask3, mask4, keep_prob, seed, sample_offset, gather=None):
    """gather = (T, slots): states / actions are a pool of trajectories of T transitions each (rows [s*T, (s+1)*T) = slot s)
    and the batch is the trajectories `slots` in that order, N = len(slots) * T transitions (dmfg_rnet_args.gather_*)."""
    from ._lib import DROPOUT_MASKS, DROPOUT_NONE, DROPOUT_PHILOX, MAX_GATHER, RnetArgs
    device = states.device
    N, d = states.shape
    P = rnet_param_count(d, n_fc3, n_fc4)
    a = RnetArgs()
    a.struct_size = C.sizeof(RnetArgs)
    a.params = _ptr(_require(params, "params", device, torch.float32, (P,)))
    a.states = _ptr(_require(states, "states", device, torch.float32, (N, d)))
    a.actions = _ptr(_require(actions, "actions", device, torch.float32, (N, d, d)))
    if gather is not None:
        T, slots = int(gather[0]), gather[1]
        if T < 1 or len(slots) > MAX_GATHER:
            raise ValueError("a gathered batch holds at most %d trajectories (got %d)" % (MAX_GATHER, len(slots)))
        if len(slots) and not (0 <= min(slots) and (max(slots) + 1) * T <= N):
            raise ValueError("gather slots outside the pool of %d rows" % N)
        a.gather_T = T
        for i, sl in enumerate(slots):
            a.gather_slots[i] = sl
        N = len(slots) * T
    a.d, a.n_fc3, a.n_fc4, a.N = d, int(n_fc3), int(n_fc4), N
    if mask3 is not None or mask4 is not None:
        a.dropout = DROPOUT_MASKS
        a.mask3 = _ptr(_require(mask3, "mask3", device, torch.uint8, (N, n_fc3)))
        a.mask4 = _ptr(_require(mask4, "mask4", device, torch.uint8, (N, n_fc4)))
    elif seed is not None:
        a.dropout = DROPOUT_PHILOX
        a.seed, a.sample_offset = int(seed) & (2 ** 64 - 1), int(sample_offset)
    else:
        a.dropout = DROPOUT_NONE
    a.keep_prob = float(keep_prob)
    return a, P


def rnet_forward(params, states, actions, n_fc3, n_fc4, *, mask3=None, mask4=None, keep_prob=0.4, seed=None,
                 sample_offset=0, out=None, gather=None):
    """r_net(state, action) for N transitions (networks.py:13-157).  states [N,d], actions [N,d,d] float32.
    Dropout: mask3/mask4 uint8 keep masks (parity), or `seed` for in-kernel Philox masks, or neither."""
    lib = _lib.load()
    a, _ = _rnet_args(params, states, actions, n_fc3, n_fc4, mask3, mask4, keep_prob, seed, sample_offset, gather)
    device = states.device
    N = a.N
    with _on(device):
        r = out if out is not None else torch.empty(N, dtype=torch.float32, device=device)
        a.rewards = _ptr(_require(r, "rewards", device, torch.float32, (N,)))
        check(lib.dmfg_rnet_forward(C.byref(a), _stream_ptr(device)))
    return r


def rnet_backward(params, states, actions, drewards, n_fc3, n_fc4, *, grad=None, accumulate=False, mask3=None,
                  mask4=None, keep_prob=0.4, seed=None, sample_offset=0, want_rewards=False, gather=None):
    """Flat gradient of sum_n drewards[n]*r[n] w.r.t. the parameters (forward recomputed in-kernel)."""
    lib = _lib.load()
    a, P = _rnet_args(params, states, actions, n_fc3, n_fc4, mask3, mask4, keep_prob, seed, sample_offset, gather)
    device = states.device
    N = a.N
    with _on(device):
        if grad is None:
            accumulate = False                               # every entry is written by the reduction
            grad = torch.empty(P, dtype=torch.float32, device=device)
        a.grad = _ptr(_require(grad, "grad", device, torch.float32, (P,)))
        a.drewards = _ptr(_require(drewards, "drewards", device, torch.float32, (N,)))
        a.accumulate = 1 if accumulate else 0
        r = None
        if want_rewards:
            r = torch.empty(N, dtype=torch.float32, device=device)
            a.rewards = _ptr(r)
        ws = _workspace(device, lib.dmfg_rnet_workspace_bytes(C.byref(a)))
        if ws is not None:
            a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_rnet_backward(C.byref(a), _stream_ptr(device)))
    return (grad, r) if want_rewards else grad


def rnet_backward_gen(params, states, actions, n_fc3, n_fc4, T, r_demo, num_demo_traj, *, layout="time_major",
                      grad=None, accumulate=False, mask3=None, mask4=None, keep_prob=0.4, seed=None, sample_offset=0,
                      want_rewards=False, local_sums=False):
    """The generated half of one reward update in one pass (ac_irl.py:390-418 with z_j = 1, T <= 16): forward,
    R_j = sum_t r[j,t], backward with the weight exp(R_j), gradient scaled by 1 / sum_j exp(R_j) and written (or
    added) to `grad`.  states [M*T, d], actions [M*T, d, d] in `layout` order; r_demo: rewards of the demonstration
    batch.  Returns (grad, loss [4] float64 device = {first+second, first, second, ln sum_j e^{R_j}}[, r_gen]).
    local_sums=True (data-parallel form): grad is the UNNORMALISED sum_j e^{R_j} dR_j/dparams of these trajectories and
    loss[0:2] = {sum_j e^{R_j}, sum r_demo}; reduce over ranks, then irl_dp_finalize."""
    from ._lib import IrlGenArgs
    lib = _lib.load()
    a, P = _rnet_args(params, states, actions, n_fc3, n_fc4, mask3, mask4, keep_prob, seed, sample_offset)
    device = states.device
    N = states.shape[0]
    M = N // int(T)
    if M * int(T) != N:
        raise ValueError("%d transitions are not a multiple of T=%d" % (N, T))
    g = IrlGenArgs()
    g.struct_size = C.sizeof(IrlGenArgs)
    g.T, g.M, g.n_demo = int(T), M, r_demo.numel()
    if layout == "time_major":
        g.gen_t_stride, g.gen_j_stride = M, 1
    elif layout == "trajectory_major":
        g.gen_t_stride, g.gen_j_stride = 1, int(T)
    else:
        raise ValueError("layout must be 'time_major' or 'trajectory_major'")
    g.num_demo_traj = float(num_demo_traj)
    g.local_sums = 1 if local_sums else 0
    g.r_demo = _ptr(_require(r_demo, "r_demo", device, torch.float32, tuple(r_demo.shape)))
    with _on(device):
        if grad is None:
            accumulate = False
            grad = torch.empty(P, dtype=torch.float32, device=device)
        a.grad = _ptr(_require(grad, "grad", device, torch.float32, (P,)))
        a.accumulate = 1 if accumulate else 0
        r = None
        if want_rewards:
            r = torch.empty(N, dtype=torch.float32, device=device)
            a.rewards = _ptr(r)
        loss = torch.empty(4, dtype=torch.float64, device=device)
        g.loss_out = _ptr(loss)
        ws = _workspace(device, lib.dmfg_rnet_workspace_bytes(C.byref(a)))
        if ws is not None:
            a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_rnet_backward_gen(C.byref(a), C.byref(g), _stream_ptr(device)))
    return (grad, loss, r) if want_rewards else (grad, loss)


def irl_dp_finalize(reduced, n):
    """After the all-reduce of a data-parallel reward step: reduced [2n+4] float64 = {grad_demo for dL/dr = -1,
    unnormalised generated gradient, Z, sum r_demo, demo trajectories, generated trajectories} summed over ranks.
    Returns (grad [n] float32 = grad_demo / N_demo + grad_gen / Z, loss [4] float64)."""
    lib = _lib.load()
    device = reduced.device
    _require(reduced, "reduced", device, torch.float64, (2 * n + 4,))
    with _on(device):
        grad = torch.empty(n, dtype=torch.float32, device=device)
        loss = torch.empty(4, dtype=torch.float64, device=device)
        check(lib.dmfg_irl_dp_finalize(int(n), _ptr(reduced), _ptr(grad), _ptr(loss), _stream_ptr(device)))
    return grad, loss


def irl_loss_grad(r_demo, r_gen, T, num_demo_traj, *, layout="time_major", log_z=None, want_grads=True):
    """IRL loss terms and dL/dr (ac_irl.py:390-406).  r_gen holds M*T rewards, time-major [T,M] (rollout
    record) or trajectory-major [M,T] (the reference's feed).  Returns dict(loss [4] float64 device =
    {first+second, first, second, ln sum_j z_j e^{R_j}}, d_demo, d_gen)."""
    from ._lib import IrlLossArgs
    lib = _lib.load()
    device = r_gen.device
    n_gen = r_gen.numel()
    M = n_gen // int(T)
    if M * int(T) != n_gen:
        raise ValueError("r_gen has %d elements, not a multiple of T=%d" % (n_gen, T))
    a = IrlLossArgs()
    a.struct_size = C.sizeof(IrlLossArgs)
    a.T, a.n_demo, a.M = int(T), r_demo.numel(), M
    if layout == "time_major":
        a.gen_t_stride, a.gen_j_stride = M, 1
    elif layout == "trajectory_major":
        a.gen_t_stride, a.gen_j_stride = 1, int(T)
    else:
        raise ValueError("layout must be 'time_major' or 'trajectory_major'")
    a.num_demo_traj = float(num_demo_traj)
    a.r_demo = _ptr(_require(r_demo, "r_demo", device, torch.float32, tuple(r_demo.shape)))
    a.r_gen = _ptr(_require(r_gen, "r_gen", device, torch.float32, tuple(r_gen.shape)))
    if log_z is not None:
        a.log_z = _ptr(_require(log_z, "log_z", device, torch.float32, (M,)))
    res = {}
    with _on(device):
        res["loss"] = torch.empty(4, dtype=torch.float64, device=device)
        a.loss_out = _ptr(res["loss"])
        if want_grads:
            res["d_demo"] = torch.empty_like(r_demo)
            res["d_gen"] = torch.empty_like(r_gen)
            a.d_demo, a.d_gen = _ptr(res["d_demo"]), _ptr(res["d_gen"])
        ws = _workspace(device, lib.dmfg_irl_loss_workspace_bytes(M))
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        check(lib.dmfg_irl_loss_grad(C.byref(a), _stream_ptr(device)))
    return res


def adam_tf(params, m, v, grad, step, lr, *, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, l1l2=False,
            net=None, want_reg_loss=False):
    """One TF-style Adam step in place (ac_irl.py:417-418); net=(d,n_fc3,n_fc4) is needed for l1l2."""
    lib = _lib.load()
    device = params.device
    n = params.numel()
    for name, t in (("params", params), ("m", m), ("v", v), ("grad", grad)):
        _require(t, name, device, torch.float32, (n,))
    d, n3, n4 = net if net is not None else (0, 0, 0)
    reg = None
    with _on(device):
        if want_reg_loss:
            reg = torch.empty(1, dtype=torch.float64, device=device)
        check(lib.dmfg_adam_tf(n, _ptr(params), _ptr(m), _ptr(v), _ptr(grad), float(grad_scale), int(step), float(lr),
                               float(beta1), float(beta2), float(eps), 1 if l1l2 else 0, int(d), int(n3), int(n4),
                               _ptr(reg), _stream_ptr(device)))
    return reg


def irl_reward_step(params, m, v, step, lr, demo_states, demo_actions, demo_weight, gen_states, gen_actions, n_fc3, n_fc4, T,
                    num_demo_traj, *, layout="time_major", keep_prob=0.4, demo_seed=None, demo_sample_offset=0,
                    gen_seed=None, gen_sample_offset=0, beta1=0.9, beta2=0.999, eps=1e-8, l1l2=False, want_reg_loss=False,
                    finishing_launch=True, demo_gather=None, gen_gather=None):
    """One whole reward update on one rank (AC_IRL.update_reward, ac_irl.py:804-846, z_j = 1) through ONE C call:
    rnet_backward(demonstrations, dL/dr = demo_weight) -> rnet_backward_gen(generated) -> adam_tf, with everything behind the
    two backward launches (reductions, loss terms, Adam) in one finishing launch; bit-identical parameters.  Returns (grad [P] float32, loss [4] float64 device, reg [1] float64 device or None)."""
    from ._lib import IrlGenArgs, IrlStepArgs
    lib = _lib.load()
    device = demo_states.device
    a, P = _rnet_args(params, demo_states, demo_actions, n_fc3, n_fc4, None, None, keep_prob, demo_seed, demo_sample_offset,
                      demo_gather)
    b, _ = _rnet_args(params, gen_states, gen_actions, n_fc3, n_fc4, None, None, keep_prob, gen_seed, gen_sample_offset,
                      gen_gather)
    n_demo, n_gen = a.N, b.N
    M = n_gen // int(T)
    if M * int(T) != n_gen:
        raise ValueError("%d transitions are not a multiple of T=%d" % (n_gen, T))
    g = IrlGenArgs()
    g.struct_size = C.sizeof(IrlGenArgs)
    g.T, g.M = int(T), M
    if layout == "time_major":
        g.gen_t_stride, g.gen_j_stride = M, 1
    elif layout == "trajectory_major":
        g.gen_t_stride, g.gen_j_stride = 1, int(T)
    else:
        raise ValueError("layout must be 'time_major' or 'trajectory_major'")
    g.num_demo_traj = float(num_demo_traj)
    st = IrlStepArgs()
    st.struct_size = C.sizeof(IrlStepArgs)
    st.l1l2 = 1 if l1l2 else 0
    for name, t in (("m", m), ("v", v)):
        _require(t, name, device, torch.float32, (P,))
    st.params, st.m, st.v = a.params, _ptr(m), _ptr(v)
    st.step, st.lr, st.beta1, st.beta2, st.eps = int(step), float(lr), float(beta1), float(beta2), float(eps)
    with _on(device):
        grad = torch.empty(P, dtype=torch.float32, device=device)
        r_demo = torch.empty(n_demo, dtype=torch.float32, device=device)
        loss = torch.empty(4, dtype=torch.float64, device=device)
        reg = torch.empty(1, dtype=torch.float64, device=device) if want_reg_loss else None
        a.grad, a.rewards = _ptr(grad), _ptr(r_demo)
        a.drewards = _ptr(_require(demo_weight, "demo_weight", device, torch.float32, (n_demo,)))
        g.loss_out = _ptr(loss)
        st.reg_loss_out = _ptr(reg)
        # (finishing_launch=False hands over the smaller workspace of one backward call: the entry point then runs the
        # chain's own six launches -- kept reachable for the parity test of the two forms)
        need = (lib.dmfg_irl_reward_step_workspace_bytes if finishing_launch else lib.dmfg_rnet_workspace_bytes)(C.byref(a))
        ws = _workspace(device, need)
        if ws is not None:
            a.workspace, a.workspace_bytes = _ptr(ws), need
        check(lib.dmfg_irl_reward_step(C.byref(a), C.byref(b), C.byref(g), C.byref(st), _stream_ptr(device)))
    return grad, loss, reg


def dirichlet_logq(states, actions, thetas, shift):
    """logq[n,k] = sum_i ln Dir(a_n[i,:]; max(alpha_{theta_k}(s_n)[i,:], 1+1e-6))  (ac_irl.py:344-361)."""
    lib = _lib.load()
    device = states.device
    N, d = states.shape
    K = thetas.numel()
    _require(states, "states", device, torch.float32, (N, d))
    _require(actions, "actions", device, torch.float32, (N, d, d))
    _require(thetas, "thetas", device, torch.float64, (K,))
    with _on(device):
        out = torch.empty((N, K), dtype=torch.float64, device=device)
        check(lib.dmfg_dirichlet_logq(d, N, K, _ptr(states), _ptr(actions), _ptr(thetas), float(shift), _ptr(out),
                                      _stream_ptr(device)))
    return out


def irl_log_z(logq, T, num_start_samples, layout="time_major"):
    """ln z_j (ac_irl.py:377-379) from per-transition logq [M*T, K]."""
    lib = _lib.load()
    device = logq.device
    NT, K = logq.shape
    M = NT // int(T)
    ts, js = (M, 1) if layout == "time_major" else (1, int(T))
    _require(logq, "logq", device, torch.float64, (NT, K))
    with _on(device):
        out = torch.empty(M, dtype=torch.float32, device=device)
        check(lib.dmfg_irl_log_z(M, int(T), K, ts, js, _ptr(logq), float(num_start_samples), _ptr(out),
                                 _stream_ptr(device)))
    return out
