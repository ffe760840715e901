"""NumPy restatement of the reference's rollout / actor-critic arithmetic.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) -- the checker for the
CUDA path and the timed CPU baseline of ``bench.py``; never imported by the
product package.

Every function cites the reference lines it follows (paths are relative to
the upstream repository ``011235813/discrete_mean_field_game``).  All noise is
an explicit input (or an explicit ``NoiseSource``), because the reference is
unseeded (``mfg_ac2.py:242,466``).  Everything broadcasts over leading batch
dimensions, so the same code serves the single-population reference semantics
and the batched ("B populations") checks; ``dtype`` lets the tests run the
identical formulas in float32 to measure cancellation.

Pinned against the reference itself by ``oracle/make_golden.py`` ->
``tests/golden/*.npz`` -> ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import special

LOG_P_ZERO = 1e-100   # mfg_ac2.py:369  P[P==0] = 1e-100
GAMMA_ZERO = 1e-20    # mfg_ac2.py:244  y[y==0] = 1e-20


def num_features(d: int) -> int:
    """F = d(d+1)/2 + d + 1 (mfg_ac2.py:175)."""
    return d * (d + 1) // 2 + d + 1


# --------------------------------------------------------------------------
# a1: policy concentration and sampling
# --------------------------------------------------------------------------
def policy_alpha(pi, theta, shift, dtype=np.float64):
    """alpha_ij = ln(1+exp(theta*(pi_j-pi_i-shift))) and d alpha/d theta.

    mfg_ac2.py:225-234 (alpha, alpha') and ac_irl.py:521-524,581-588.
    The naive ``log(1+exp(.))`` is kept on purpose: it is the reference's
    arithmetic in float64.  ``theta``/``shift`` may be arrays broadcastable to
    the batch shape (independent learners).
    """
    pi = np.asarray(pi, dtype=dtype)
    theta = np.asarray(theta, dtype=dtype)[..., None, None]
    shift = np.asarray(shift, dtype=dtype)[..., None, None]
    x = pi[..., None, :] - pi[..., :, None] - shift          # x[i,j] = pi_j - pi_i - shift
    with np.errstate(over="ignore"):
        alpha = np.log(1 + np.exp(theta * x))
        alpha_deriv = x / (1 + np.exp(-theta * x))
    return alpha, alpha_deriv


def normalise_gamma(y, dtype=np.float64):
    """Rows of Gamma variates -> rows on the simplex (mfg_ac2.py:244-249)."""
    y = np.array(y, dtype=dtype, copy=True)
    y[y == 0] = GAMMA_ZERO
    return y / y.sum(axis=-1, keepdims=True)


def sample_action(pi, theta, shift, alpha_scale, y=None, dtype=np.float64):
    """P for one or many populations; returns (P, alpha, alpha_deriv, y).

    With ``y is None`` the draw pattern of mfg_ac2.py:238-242 is reproduced
    exactly -- one ``np.random.gamma(shape=row)`` call per row, in row order,
    on the *global* NumPy state -- so that ``np.random.seed(s)`` yields the
    reference's own stream (only valid for a single population).
    """
    alpha, alpha_deriv = policy_alpha(pi, theta, shift, dtype)
    if y is None:
        if alpha.ndim != 2:
            raise ValueError("global-RNG sampling follows the reference and is single-population")
        d = alpha.shape[-1]
        y = np.empty_like(alpha)
        for i in range(d):
            y[i] = np.random.gamma(shape=alpha[i, :] * alpha_scale, scale=1)
    P = normalise_gamma(y, dtype)
    return P, alpha, alpha_deriv, np.asarray(y, dtype=dtype)


# --------------------------------------------------------------------------
# a2/a3: dynamics and rewards
# --------------------------------------------------------------------------
def mean_field_step(P, pi):
    """pi'_j = sum_i pi_i P_ij (mfg_ac2.py:497; ac_irl.py:679,763)."""
    return np.einsum("...i,...ij->...j", pi, P)


def reward_ac2(P, pi):
    """r = <pi, (P*P) pi - ((P*P) 1) * pi>  (mfg_ac2.py:274-279)."""
    P2 = P * P
    v1 = np.einsum("...ij,...j->...i", P2, pi)
    v2 = P2.sum(axis=-1) * pi
    return np.einsum("...i,...i->...", pi, v1 - v2)


def reward_synthetic(P, pi):
    """r = -1/2 sum_i pi_i ||P_i||^2  (mfg_synthetic.py:262-263)."""
    return -0.5 * np.einsum("...i,...i->...", pi, (P * P).sum(axis=-1))


def cost_ac1(P, pi):
    """mfg_ac.py:206-211 -- the v1 cost is minus the v2 reward."""
    return -reward_ac2(P, pi)


REWARDS = {"ac2": reward_ac2, "synthetic": reward_synthetic}


# --------------------------------------------------------------------------
# a4: critic features
# --------------------------------------------------------------------------
def feature_index_pairs(d: int):
    """(i, j) for i <= j in itertools.combinations_with_replacement order."""
    iu = np.triu_indices(d)
    return iu[0], iu[1]


def features(pi):
    """phi(pi) = [pi_i pi_j (i<=j, row-major), pi_0..pi_{d-1}, 1].

    Code order of mfg_ac2.py:333-344 (the docstring at :169-171 lists a
    different order; the code is what runs -- SURVEY App. B).
    """
    pi = np.asarray(pi)
    d = pi.shape[-1]
    i, j = feature_index_pairs(d)
    quad = pi[..., i] * pi[..., j]
    one = np.ones(pi.shape[:-1] + (1,), dtype=pi.dtype)
    return np.concatenate([quad, pi, one], axis=-1)


def value(pi, w):
    """V(pi; w) = phi(pi) . w  (mfg_ac2.py:290-311)."""
    return features(pi) @ np.asarray(w).reshape(-1)


# --------------------------------------------------------------------------
# a6: policy-gradient scalar
# --------------------------------------------------------------------------
def log_policy_gradient(alpha, alpha_deriv, P):
    """g = sum_ij (-psi(alpha_ij) + psi(sum_j alpha_ij) + ln P_ij) alpha'_ij.

    mfg_ac2.py:364-379 / ac_irl.py:608-622.  Uses the *unscaled* alpha
    (quirk C.1) and P==0 -> 1e-100; unlike the reference it does not mutate P.
    """
    P = np.where(P == 0, np.asarray(LOG_P_ZERO, dtype=np.float64), P)
    term = (-special.digamma(alpha)
            + special.digamma(alpha.sum(axis=-1, keepdims=True))
            + np.log(P))
    return (term * alpha_deriv).sum(axis=(-1, -2))


# --------------------------------------------------------------------------
# a5/a7: step sizes
# --------------------------------------------------------------------------
def critic_lr(episode, lr_critic, constant):
    """mfg_ac2.py:511-514."""
    return lr_critic if constant else lr_critic / (episode + 1)


def actor_lr(episode, lr_actor, constant):
    """mfg_ac2.py:519-522."""
    return lr_actor if constant else lr_actor / ((episode + 1) * math.log(math.log(episode + 20)))


# --------------------------------------------------------------------------
# one full transition with frozen parameters (a1..a6), batched
# --------------------------------------------------------------------------
def transition(pi, theta, shift, alpha_scale, y, w=None, gamma_next=1.0,
               reward="ac2", dtype=np.float64):
    """One population-step.  Returns a dict of every intermediate.

    ``gamma_next`` multiplies V(pi'): gamma for mfg_ac2.py:505, the cumulative
    discount gamma^t for ac_irl.py:691.  ``reward`` is a key of REWARDS, or an
    array of externally supplied rewards (the IRL reward net).
    """
    P, alpha, alpha_deriv, y = sample_action(pi, theta, shift, alpha_scale, y, dtype)
    pi = np.asarray(pi, dtype=dtype)
    pi_next = mean_field_step(P, pi)
    if isinstance(reward, str):
        r = REWARDS[reward](P, pi) if reward != "none" else np.zeros(pi.shape[:-1], dtype)
    else:
        r = np.asarray(reward, dtype=dtype)
    g = log_policy_gradient(alpha, alpha_deriv, P).astype(dtype)
    out = dict(P=P, alpha=alpha, alpha_deriv=alpha_deriv, pi_next=pi_next, reward=r, grad=g)
    if w is not None:
        w = np.asarray(w, dtype=dtype)
        phi = features(pi)
        phi_next = features(pi_next)
        if w.ndim == 1:
            v, v_next = phi @ w, phi_next @ w
        else:                                   # per-learner weights [..., F]
            v, v_next = (phi * w).sum(-1), (phi_next * w).sum(-1)
        out.update(phi=phi, v=v, v_next=v_next, delta=r + gamma_next * v_next - v)
    return out


def rollout_frozen(pi0, theta, shift, alpha_scale, y, w=None, gamma=1.0,
                   discount="step", reward="ac2", dtype=np.float64):
    """T transitions for B populations with frozen (theta, w).

    pi0 [B,d]; y [T,B,d,d] (time-major, the device layout).  Returns
    time-major arrays: states [T+1,B,d], actions [T,B,d,d], alpha,
    alpha_deriv, rewards/deltas/grads [T,B] and the per-episode sums
    G_theta = sum delta*g, G_w = sum delta*phi, R = sum r that the
    ``per_episode`` update consumes.
    """
    pi = np.asarray(pi0, dtype=dtype)
    T = y.shape[0]
    B, d = pi.shape
    F = num_features(d)
    states = np.zeros((T + 1, B, d), dtype)
    actions = np.zeros((T, B, d, d), dtype)
    alphas = np.zeros((T, B, d, d), dtype)
    derivs = np.zeros((T, B, d, d), dtype)
    rewards = np.zeros((T, B), dtype)
    deltas = np.zeros((T, B), dtype)
    grads = np.zeros((T, B), dtype)
    G_theta, G_w, R = 0.0, np.zeros(F), 0.0
    states[0] = pi
    disc = 1.0
    for t in range(T):
        g_next = gamma if discount == "step" else disc   # ac_irl.py:691 uses the running product
        rew = reward if isinstance(reward, str) else reward[t]
        o = transition(pi, theta, shift, alpha_scale, y[t], w, g_next, rew, dtype)
        actions[t], alphas[t], derivs[t] = o["P"], o["alpha"], o["alpha_deriv"]
        rewards[t], grads[t] = o["reward"], o["grad"]
        if w is not None:
            deltas[t] = o["delta"]
            G_theta += float((o["delta"].astype(np.float64) * o["grad"]).sum())
            G_w += (o["delta"][:, None].astype(np.float64) * o["phi"]).sum(axis=0)
        R += float(o["reward"].sum())
        disc *= gamma
        pi = o["pi_next"]
        states[t + 1] = pi
    return dict(states=states, actions=actions, alpha=alphas, alpha_deriv=derivs,
                rewards=rewards, deltas=deltas, grads=grads,
                G_theta=G_theta, G_w=G_w, R=R)


# --------------------------------------------------------------------------
# a8: the serial learner -- the reference's train() loop, noise made explicit
# --------------------------------------------------------------------------
class GlobalNumpyNoise:
    """The reference's own noise: the process-global NumPy RNG.

    Call order matches mfg_ac2.py:466 (randint) and :242 (one gamma call per
    row), so seeding NumPy reproduces a reference run draw for draw.
    """

    def start_index(self, n):
        return int(np.random.randint(n))

    def gamma_rows(self, shape_matrix):
        y = np.empty_like(shape_matrix)
        for i in range(shape_matrix.shape[0]):
            y[i] = np.random.gamma(shape=shape_matrix[i, :], scale=1)
        return y


@dataclass
class InjectedNoise:
    """Replays recorded start rows [E] and Gamma variates [E,T,d,d]."""
    start_rows: np.ndarray
    y: np.ndarray
    _e: int = field(default=-1)
    _t: int = field(default=0)

    def start_index(self, n):
        self._e += 1
        self._t = 0
        return int(self.start_rows[self._e])

    def gamma_rows(self, shape_matrix):
        y = self.y[self._e, self._t]
        self._t += 1
        return y


def train_serial(mat_pi0, theta, w, shift, alpha_scale, num_episodes, *, gamma=1.0,
                 constant=False, lr_critic=0.1, lr_actor=0.001, flavour="mfg_ac2",
                 reward="ac2", reward_fn=None, noise=None, num_steps=15, trace=False):
    """One learner, per-step online updates: mfg_ac2.py:448-539 / ac_irl.py:634-732.

    ``flavour`` selects the two places where the files differ (quirks C.5, C.6):
    episode index origin (0 vs 1) and the factor on V(pi') (gamma vs the
    running discount).  ``reward_fn(pi, P) -> float`` replaces the closed form
    (the reward-net query of ac_irl.py:683).  Returns (theta, w, info).
    """
    noise = noise or GlobalNumpyNoise()
    w = np.array(w, dtype=np.float64).reshape(-1)
    theta = float(theta)
    first = 0 if flavour == "mfg_ac2" else 1
    info = dict(total_reward=[], theta=[], pi_final=[])
    steps = [] if trace else None
    for episode in range(first, first + num_episodes):
        pi = np.array(mat_pi0[noise.start_index(mat_pi0.shape[0])], dtype=np.float64)
        discount, total = 1.0, 0.0
        for _ in range(num_steps):
            alpha, alpha_deriv = policy_alpha(pi, theta, shift)
            y = np.asarray(noise.gamma_rows(alpha * alpha_scale), dtype=np.float64)
            P = normalise_gamma(y)
            pi_next = mean_field_step(P, pi)
            r = float(reward_fn(pi, P)) if reward_fn is not None else float(REWARDS[reward](P, pi))
            phi, phi_next = features(pi), features(pi_next)
            g_next = gamma if flavour == "mfg_ac2" else discount
            delta = r + g_next * float(phi_next @ w) - float(phi @ w)
            g = float(log_policy_gradient(alpha, alpha_deriv, P))
            w_new = w + critic_lr(episode, lr_critic, constant) * delta * phi
            theta_new = theta + actor_lr(episode, lr_actor, constant) * delta * g
            if trace:
                steps.append(dict(episode=episode, pi=pi, y=y, P=P, pi_next=pi_next, reward=r,
                                  delta=delta, grad=g, theta=theta_new, w=w_new.copy()))
            w, theta = w_new, theta_new
            discount *= gamma
            pi = pi_next
            total += r
        info["total_reward"].append(total)
        info["theta"].append(theta)
        info["pi_final"].append(pi)
    if trace:
        info["steps"] = steps
    return theta, w, info


def generate_trajectory(pi0, total_hours, theta, shift, alpha_scale, noise=None):
    """[total_hours, d] states, rollout only (mfg_ac2.py:566-592)."""
    noise = noise or GlobalNumpyNoise()
    pi = np.asarray(pi0, dtype=np.float64)
    out = np.zeros((total_hours, pi.shape[0]))
    out[0] = pi
    for h in range(1, total_hours):
        alpha, _ = policy_alpha(pi, theta, shift)
        P = normalise_gamma(noise.gamma_rows(alpha * alpha_scale))
        pi = mean_field_step(P, pi)
        out[h] = pi
    return out


# --------------------------------------------------------------------------
# evaluation metrics of generate_trajectory's consumer (mfg_ac2.py:546-563, 595-670)
# --------------------------------------------------------------------------
def jsd(P, Q):
    """Jensen-Shannon divergence as mfg_ac2.py:546-563 computes it: zeros -> 1e-100, M = (P+Q)/2 from the
    UNNORMALISED inputs, then scipy.stats.entropy semantics (each argument normalised to sum 1)."""
    P = np.where(np.asarray(P, dtype=np.float64) == 0, 1e-100, P)
    Q = np.where(np.asarray(Q, dtype=np.float64) == 0, 1e-100, Q)
    M = 0.5 * (P + Q)

    def kl(a, b):
        a = a / a.sum(-1, keepdims=True)
        b = b / b.sum(-1, keepdims=True)
        return (a * np.log(a / b)).sum(-1)
    return 0.5 * (kl(P, M) + kl(Q, M))


def trajectory_metrics(generated, empirical):
    """Per-trajectory (l1_final, l1_mean, JSD_final, JSD_mean) of mfg_ac2.py:627-650.
    generated / empirical: [B, H, d]."""
    g = np.asarray(generated, dtype=np.float64)
    e = np.asarray(empirical, dtype=np.float64)
    l1 = np.abs(e - g).sum(-1)                      # [B, H]
    js = jsd(g, e)                                  # [B, H]
    return l1[:, -1], l1.mean(1), js[:, -1], js.mean(1)


def jsd_synthetic(P, Q):
    """mfg_synthetic.py:529-546: like jsd() but every entry <= 0 (not only == 0) becomes 1e-100 -- the analytic
    rows V_j - V_i do go negative."""
    P = np.where(np.asarray(P, dtype=np.float64) <= 0, 1e-100, P)
    Q = np.where(np.asarray(Q, dtype=np.float64) <= 0, 1e-100, Q)
    M = 0.5 * (P + Q)

    def kl(a, b):
        a = a / a.sum(-1, keepdims=True)
        b = b / b.sum(-1, keepdims=True)
        return (a * np.log(a / b)).sum(-1)
    return 0.5 * (kl(P, M) + kl(Q, M))


def synthetic_check(actions):
    """The analytic check of mfg_synthetic.py:741-899 for ONE trajectory.  actions [T, d, d].
    Backward equation V^n = r(P^n) + P^n V^{n+1}, V^T = 0, r_i = -1/2 ||P_i||^2 (:726-738, :775-778); then per
    step n the matrix the MFG theory predicts, A_ij = V_j - V_i (i != j), A_ii = 1 - (sum_j V_j - d V_i)
    (:786-794), and  l1[n] = sum_ij |P_ij - A_ij|  (:795),  jsd[n] = sum_i JSD(P_i, A_i)  (:866-877)."""
    A = np.asarray(actions, dtype=np.float64)
    T, d, _ = A.shape
    V = np.zeros(d)
    l1, js = np.zeros(T), np.zeros(T)
    for n in range(T - 1, -1, -1):
        V = -0.5 * (A[n] ** 2).sum(1) + A[n] @ V
        pred = V[None, :] - V[:, None]
        pred[np.arange(d), np.arange(d)] = 1.0 - (V.sum() - d * V)
        l1[n] = np.abs(A[n] - pred).sum()
        js[n] = jsd_synthetic(A[n], pred).sum()
    return l1, js


def synthetic_start_states(n_rows=21, n_cols=20, d=15, seed=0):
    """Dirichlet(1_20) rows rounded to %.3e, first d columns,
    not renormalised (mirrors the parse at mfg_ac2.py:191-198)."""
    rng = np.random.RandomState(seed)
    rows = rng.dirichlet(np.ones(n_cols), size=n_rows)
    rows = np.array([[float("%.3e" % v) for v in row] for row in rows])
    return rows[:, :d].copy()


# --------------------------------------------------------------------------
# the batched ("per-episode batch-mean") trainer of the GPU path, restated on the CPU with NumPy noise:
# what bench.py --impl reference times when it is asked for the SAME workload as the GPU arm
# --------------------------------------------------------------------------
def train_batch_port(pi0, theta, w, shift, alpha_scale, num_episodes, *, T=16, gamma=1.0, lr_critic=0.1,
                     lr_actor=0.1, first_episode=0, reward="ac2", rng=None):
    """B populations share (theta, w), frozen within an episode; after every episode
    theta += lr_a(e)/B sum delta*g, w += lr_c(e)/B sum delta*phi (the reference's updates, mfg_ac2.py:505-522,
    applied to the batch mean).  Gamma variates from ``rng`` (vectorised np.random.gamma).  Returns (theta, w)."""
    rng = rng or np.random
    pi0 = np.asarray(pi0, dtype=np.float64)
    B = pi0.shape[0]
    w = np.array(w, dtype=np.float64).reshape(-1)
    theta = float(theta)
    for e in range(num_episodes):
        episode = first_episode + e
        pi = pi0
        G_theta, G_w = 0.0, np.zeros_like(w)
        for _ in range(T):
            alpha, _ = policy_alpha(pi, theta, shift)
            y = rng.gamma(shape=alpha * alpha_scale, scale=1.0)
            o = transition(pi, theta, shift, alpha_scale, y, w, gamma, reward)
            G_theta += float((o["delta"] * o["grad"]).sum())
            G_w += (o["delta"][:, None] * o["phi"]).sum(axis=0)
            pi = o["pi_next"]
        theta += actor_lr(episode, lr_actor, False) / B * G_theta
        w = w + critic_lr(episode, lr_critic, False) / B * G_w
    return theta, w
