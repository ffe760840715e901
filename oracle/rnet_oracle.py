"""NumPy float64 restatement of the reference's reward network, IRL loss, TF-Adam and calc_z.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) -- the checker for the CUDA kernels and the
timed CPU baseline of the IRL path; never imported by the product package.

PARITY UNPINNED at the TensorFlow boundary: ``networks.py`` / ``ac_irl.py`` need TensorFlow 1.x
(``tf.contrib``), which cannot be installed here, and the reference's ``test_acirl.py`` records no
expected values.  This file restates the documented semantics of the TF ops the reference calls
(file:line cited per function); ``tests/test_rnet_oracle.py`` cross-checks it against an independent
``torch.nn.functional`` + autograd implementation, finite differences and the hand-rolled Dirichlet
pdf of ``test_acirl.py:15-30``.

Parameters live in ONE flat vector in TF variable-creation order (``layout``), the layout the device
kernels use.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import special

F1, K1, F2, K2 = 1, 5, 2, 3          # ac_irl.py:249-267 always passes f1=1, k1=5, f2=2, k2=3
KEEP_PROB = 0.4                       # networks.py:70,75
REG_KINDS = ("none", "dropout", "l1l2", "dropout_l1l2")   # ac_irl.py:35


# --------------------------------------------------------------------------- parameters
def layout(d, n_fc3, n_fc4, f1=F1, k1=K1, f2=F2, k2=K2):
    """[(name, shape, offset)] of the flat parameter vector, TF variable order
    (networks.py:27-41: conv1, conv2, fc3, fc4, out; weights then biases)."""
    shapes = [("conv1/weights", (k1, k1, 1, f1)), ("conv1/biases", (f1,)),
              ("conv2/weights", (k2, k2, f1, f2)), ("conv2/biases", (f2,)),
              ("fc3/weights", (f2 * d * d, n_fc3)), ("fc3/biases", (n_fc3,)),
              ("fc4/weights", (n_fc3 + d, n_fc4)), ("fc4/biases", (n_fc4,)),
              ("out/weights", (n_fc4, 1)), ("out/biases", (1,))]
    out, off = [], 0
    for name, shp in shapes:
        out.append((name, shp, off))
        off += int(np.prod(shp))
    return out


def param_count(d, n_fc3, n_fc4):
    name, shp, off = layout(d, n_fc3, n_fc4)[-1]
    return off + int(np.prod(shp))


def unpack(params, d, n_fc3, n_fc4):
    params = np.asarray(params)
    return {name: params[off:off + int(np.prod(shp))].reshape(shp) for name, shp, off in layout(d, n_fc3, n_fc4)}


def xavier_init(d, n_fc3, n_fc4, rng):
    """tf.contrib.layers default initialisers [3p]: xavier_initializer(uniform=True) for weights
    (limit = sqrt(6/(fan_in+fan_out)), conv fans include the receptive field), zeros for biases."""
    p = np.zeros(param_count(d, n_fc3, n_fc4))
    for name, shp, off in layout(d, n_fc3, n_fc4):
        if name.endswith("biases"):
            continue
        if len(shp) == 4:
            rf = shp[0] * shp[1]
            fan_in, fan_out = rf * shp[2], rf * shp[3]
        else:
            fan_in, fan_out = shp
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        n = int(np.prod(shp))
        p[off:off + n] = rng.uniform(-lim, lim, size=n)
    return p


# --------------------------------------------------------------------------- forward / backward
def _conv_same(x, w):
    """NHWC cross-correlation, stride 1, SAME padding (tf.contrib.layers.conv2d, networks.py:27,30).
    x [B,H,W,Cin], w [k,k,Cin,Cout] -> [B,H,W,Cout]."""
    k = w.shape[0]
    p = k // 2
    B, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)))
    out = np.zeros((B, H, W, w.shape[3]), dtype=x.dtype)
    for dh in range(k):
        for dw in range(k):
            out += np.einsum("bhwc,co->bhwo", xp[:, dh:dh + H, dw:dw + W, :], w[dh, dw])
    return out


def _conv_same_grad_w(x, gout, k):
    """d/dw of _conv_same: [k,k,Cin,Cout]."""
    p = k // 2
    B, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)))
    gw = np.zeros((k, k, x.shape[3], gout.shape[3]), dtype=x.dtype)
    for dh in range(k):
        for dw in range(k):
            gw[dh, dw] = np.einsum("bhwc,bhwo->co", xp[:, dh:dh + H, dw:dw + W, :], gout)
    return gw


def _conv_same_grad_x(gout, w):
    """d/dx of _conv_same: [B,H,W,Cin]."""
    k = w.shape[0]
    p = k // 2
    B, H, W, _ = gout.shape
    gp = np.pad(gout, ((0, 0), (p, p), (p, p), (0, 0)))
    gx = np.zeros((B, H, W, w.shape[2]), dtype=gout.dtype)
    for dh in range(k):
        for dw in range(k):
            # out[h,w] += x[h+dh-p, w+dw-p] w[dh,dw]  =>  gx[h',w'] += gout[h'-dh+p, w'-dw+p] w[dh,dw]
            gx += np.einsum("bhwo,co->bhwc", gp[:, 2 * p - dh:2 * p - dh + H, 2 * p - dw:2 * p - dw + W, :], w[dh, dw])
    return gx


def forward(params, states, actions, n_fc3, n_fc4, mask3=None, mask4=None, keep_prob=KEEP_PROB, cache=False):
    """r = r_net(state, action) for N transitions (networks.py:13-43; dropout variants :46-157).

    states [N,d], actions [N,d,d].  mask3 [N,n_fc3] / mask4 [N,n_fc4] are 0/1 keep masks of the two
    dropout layers (None = no dropout, i.e. reg in {'none','l1l2'}); kept units are scaled by
    1/keep_prob (tf.contrib.layers.dropout with its default is_training=True [3p])."""
    states = np.asarray(states, dtype=np.float64)
    actions = np.asarray(actions, dtype=np.float64)
    N, d = states.shape
    P = unpack(np.asarray(params, dtype=np.float64), d, n_fc3, n_fc4)
    a = actions.reshape(N, d, d, 1)                                              # networks.py:23
    z1 = _conv_same(a, P["conv1/weights"]) + P["conv1/biases"]
    c1 = np.maximum(z1, 0)                                                       # :27 relu
    z2 = _conv_same(c1, P["conv2/weights"]) + P["conv2/biases"]
    c2 = np.maximum(z2, 0)                                                       # :30
    flat = c2.reshape(N, F2 * d * d)                                             # :32 NHWC flatten, channel fastest
    z3 = flat @ P["fc3/weights"] + P["fc3/biases"]
    h3 = np.maximum(z3, 0)                                                       # :34
    if mask3 is not None:
        h3 = h3 * np.asarray(mask3, dtype=np.float64) / keep_prob                # :70
    cat = np.concatenate([h3, states], axis=1)                                   # :36
    z4 = cat @ P["fc4/weights"] + P["fc4/biases"]
    h4 = np.maximum(z4, 0)                                                       # :38
    if mask4 is not None:
        h4 = h4 * np.asarray(mask4, dtype=np.float64) / keep_prob                # :75
    z5 = h4 @ P["out/weights"] + P["out/biases"]
    r = np.tanh(z5)[:, 0]                                                        # :41 (tanh, quirk C.10)
    if cache:
        return r, dict(a=a, z1=z1, c1=c1, z2=z2, flat=flat, z3=z3, h3=h3, cat=cat, z4=z4, h4=h4, r=r,
                       mask3=mask3, mask4=mask4, keep_prob=keep_prob, P=P, d=d, n3=n_fc3, n4=n_fc4)
    return r


def backward(cache, dr):
    """Flat gradient of sum_n dr[n] * r[n] with respect to the parameters."""
    c = cache
    P, d, n3, n4 = c["P"], c["d"], c["n3"], c["n4"]
    N = c["r"].shape[0]
    g = {}
    dz5 = (np.asarray(dr, dtype=np.float64) * (1 - c["r"] ** 2))[:, None]
    g["out/weights"] = c["h4"].T @ dz5
    g["out/biases"] = dz5.sum(0)
    dh4 = dz5 @ P["out/weights"].T
    if c["mask4"] is not None:
        dh4 = dh4 * np.asarray(c["mask4"], dtype=np.float64) / c["keep_prob"]
    dz4 = dh4 * (c["z4"] > 0)
    g["fc4/weights"] = c["cat"].T @ dz4
    g["fc4/biases"] = dz4.sum(0)
    dh3 = (dz4 @ P["fc4/weights"].T)[:, :n3]
    if c["mask3"] is not None:
        dh3 = dh3 * np.asarray(c["mask3"], dtype=np.float64) / c["keep_prob"]
    dz3 = dh3 * (c["z3"] > 0)
    g["fc3/weights"] = c["flat"].T @ dz3
    g["fc3/biases"] = dz3.sum(0)
    dc2 = (dz3 @ P["fc3/weights"].T).reshape(N, d, d, F2)
    dz2 = dc2 * (c["z2"] > 0)
    g["conv2/weights"] = _conv_same_grad_w(c["c1"], dz2, K2)
    g["conv2/biases"] = dz2.sum((0, 1, 2))
    dc1 = _conv_same_grad_x(dz2, P["conv2/weights"])
    dz1 = dc1 * (c["z1"] > 0)
    g["conv1/weights"] = _conv_same_grad_w(c["a"], dz1, K1)
    g["conv1/biases"] = dz1.sum((0, 1, 2))
    return np.concatenate([g[name].reshape(-1) for name, _, _ in layout(d, n3, n4)])


# --------------------------------------------------------------------------- loss (a11)
def reg_mask(d, n_fc3, n_fc4):
    """1 on the entries l1_l2_regularizer applies to: fc3 and fc4 WEIGHTS only
    (networks.py:69,74,110,114)."""
    m = np.zeros(param_count(d, n_fc3, n_fc4))
    for name, shp, off in layout(d, n_fc3, n_fc4):
        if name in ("fc3/weights", "fc4/weights"):
            m[off:off + int(np.prod(shp))] = 1
    return m


def reg_loss(params, d, n_fc3, n_fc4):
    """tf.contrib.layers.l1_l2_regularizer() defaults scale_l1 = scale_l2 = 1.0 [3p]:
    sum|w| + sum(w^2)/2 over fc3 and fc4 weights (ac_irl.py:409-411)."""
    w = np.asarray(params, dtype=np.float64) * reg_mask(d, n_fc3, n_fc4)
    return np.abs(w).sum() + 0.5 * (w * w).sum()


def reg_grad(params, d, n_fc3, n_fc4):
    w = np.asarray(params, dtype=np.float64)
    return (np.sign(w) + w) * reg_mask(d, n_fc3, n_fc4)


def irl_loss(r_demo, r_gen, num_demo_traj, log_z=None):
    """loss terms of ac_irl.py:390-406 and their derivatives.

    r_demo [N_transitions]; r_gen [M, T] (trajectory-major as the reference reshapes it, :397).
    first  = -(1/num_demo_traj) * sum(r_demo)      (quirk C.9: divides by trajectories, not transitions)
    second = log((1/M) sum_j z_j exp(sum_t r_gen[j,t]))   (z_j = 1 in the active code, :406)
    Returns (first, second, d first/d r_demo [N], d second/d r_gen [M,T]).
    """
    r_demo = np.asarray(r_demo, dtype=np.float64)
    r_gen = np.asarray(r_gen, dtype=np.float64)
    M = r_gen.shape[0]
    first = -r_demo.sum() / num_demo_traj
    R = r_gen.sum(axis=1)
    if log_z is not None:
        R = R + np.asarray(log_z, dtype=np.float64)
    lse = special.logsumexp(R)             # the reference exponentiates directly (:401); same value, no overflow
    second = lse - math.log(M)
    sm = np.exp(R - lse)
    return first, second, np.full(r_demo.shape, -1.0 / num_demo_traj), np.repeat(sm[:, None], r_gen.shape[1], axis=1)


def loss_and_grad(params, demo_states, demo_actions, gen_states, gen_actions, n_fc3, n_fc4, num_demo_traj, T,
                  reg="none", masks=None, log_z=None):
    """Full update_reward objective (ac_irl.py:804-846): loss, (first, second), flat gradient.
    gen_* are trajectory-major [M*T, ...]; masks = dict(demo3, demo4, gen3, gen4) for dropout variants."""
    d = np.asarray(demo_states).shape[1]
    masks = masks or {}
    rd, cd = forward(params, demo_states, demo_actions, n_fc3, n_fc4, masks.get("demo3"), masks.get("demo4"), cache=True)
    rg, cg = forward(params, gen_states, gen_actions, n_fc3, n_fc4, masks.get("gen3"), masks.get("gen4"), cache=True)
    first, second, d_demo, d_gen = irl_loss(rd, rg.reshape(-1, T), num_demo_traj, log_z)
    grad = backward(cd, d_demo) + backward(cg, d_gen.reshape(-1))
    loss = first + second
    if reg in ("l1l2", "dropout_l1l2"):
        loss += reg_loss(params, d, n_fc3, n_fc4)
        grad = grad + reg_grad(params, d, n_fc3, n_fc4)
    return loss, (first, second), grad


# --------------------------------------------------------------------------- TF-style Adam
def adam_tf(params, m, v, grad, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (ac_irl.py:417) [3p]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m,v EMA; p -= lr_t * m / (sqrt(v) + eps) -- eps OUTSIDE the bias correction.  step counts from 1."""
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    return params - lr_t * m / (np.sqrt(v) + eps), m, v


# --------------------------------------------------------------------------- calc_z (a13)
def dirichlet_logpdf_rows(P, alpha):
    """log Dir(P_i ; alpha_i) per row: lgamma(sum a) - sum lgamma(a) + sum (a-1) ln P."""
    return (special.gammaln(alpha.sum(-1)) - special.gammaln(alpha).sum(-1)
            + ((alpha - 1) * np.log(P)).sum(-1))


def log_q(states, actions, thetas, shift):
    """log prod_t prod_i Dir(a_t[i,:]; max(alpha_k(s_t)[i,:], 1+1e-6)) for every (trajectory, policy).

    ac_irl.py:332-375 without the `c` normaliser: states [M,T,d], actions [M,T,d,d], thetas [K] -> [M,K].
    NOTE the unscaled alpha and the lower clamp: q_k is not the sampling density."""
    s = np.asarray(states, dtype=np.float64)
    a = np.asarray(actions, dtype=np.float64)
    th = np.asarray(thetas, dtype=np.float64)
    diff = s[:, None, :, None, :] - s[:, None, :, :, None]                        # [M,1,T,d(i),d(j)] = s_j - s_i
    alpha = np.log(1 + np.exp((diff - shift) * th[None, :, None, None, None]))    # :353-354
    alpha = np.maximum(alpha, 1 + 1e-6)                                           # :357
    lp = dirichlet_logpdf_rows(a[:, None], alpha)                                 # [M,K,T,d]
    return lp.sum(axis=(2, 3))


def log_z(states, actions, thetas, shift, num_start_samples):
    """ln z_j, z_j = K / (N_start * sum_k q_k(tau_j))  (ac_irl.py:377-379), evaluated in log space
    (the reference divides every row pdf by `c` to stay inside float64; the result is c-free up to
    the factor c^(15 d) which cancels in the softmax weights of the loss gradient)."""
    lq = log_q(states, actions, thetas, shift)
    K = lq.shape[1]
    return math.log(K) - math.log(num_start_samples) - special.logsumexp(lq, axis=1)
