"""BASELINE.json's full sizes (configs[2]: 2^20 populations x d = 15 x 16 steps) through size-independent
properties -- the oracle cannot run these sizes, so the checks are the domain's invariants:

  * additivity: the gradient buffer of a batch is the sum of the buffers of its shards (the populations are
    independent and the Philox counters carry the GLOBAL population id), which is also what makes the
    multi-GPU sharding exact;
  * determinism: two launches with the same seed are bit-identical, whatever the grid;
  * rows of P on the simplex, mass conservation, state_{t+1} = action_t^T state_t (test2.py:26,32;
    test_acirl.py:43-47) on a recorded 2^18-population rollout (3.8 GB of actions);
  * the TD recursion: sum_t delta_t telescopes to sum_t r_t + V(pi_T) - V(pi_0) at gamma = 1
    (mfg_ac2.py:505), tying rewards, deltas, the critic evaluation and the final state together.
"""
import numpy as np
import pytest
import torch

from oracle import mfg_oracle as O

pytestmark = pytest.mark.gpu

eng = pytest.importorskip("discrete_mean_field_game_b200.engine")

D, T = 15, 16
THETA, SHIFT, SCALE = 8.86349, 0.16, 12000.0


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    eng.require_cuda()
    return torch.device("cuda:0")


def _pi0(B, dev, seed=3):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = -torch.log(torch.rand((B, D), generator=g, dtype=torch.float64))     # Exp(1) -> Dirichlet(1)
    return (x / x.sum(1, keepdim=True)).to(torch.float32).to(dev)


def _w(dev):
    return torch.as_tensor(np.random.RandomState(0).rand(O.num_features(D)), dtype=torch.float64, device=dev)


def _value(pi, w):
    """V(pi) = phi(pi) . w in double (mfg_ac2.py:290-344; features in combinations_with_replacement order)."""
    pi = pi.double()
    iu = torch.triu_indices(D, D, device=pi.device)
    W = torch.zeros((D, D), dtype=torch.float64, device=pi.device)
    W[iu[0], iu[1]] = w[:D * (D + 1) // 2]
    Q = D * (D + 1) // 2
    return torch.einsum("bi,ij,bj->b", pi, W, pi) + pi @ w[Q:Q + D] + w[Q + D]


def test_train_step_is_additive_over_shards_and_deterministic(dev):
    B = 1 << 20
    pi0, w = _pi0(B, dev), _w(dev)
    kw = dict(w=w, seed=1234, step_offset=32, outputs=(), want_acc=True)
    full = eng.rollout(pi0, THETA, SHIFT, SCALE, T, **kw)["acc"].clone()
    again = eng.rollout(pi0, THETA, SHIFT, SCALE, T, **kw)["acc"].clone()
    assert torch.equal(full, again)
    parts = torch.zeros_like(full)
    cuts = [0, 300001, 1 << 19, B - 5, B]                         # ragged shards, not multiples of the 16-population tile
    for a, b in zip(cuts[:-1], cuts[1:]):
        parts += eng.rollout(pi0[a:b].contiguous(), THETA, SHIFT, SCALE, T, pop_offset=a, **kw)["acc"]
    # the shard sums differ from the one-launch sum only by the order of ~1e6 double additions
    scale = eng.rollout(pi0[:4096].contiguous(), THETA, SHIFT, SCALE, T, w=w, seed=1234, step_offset=32,
                        outputs=("deltas",))["deltas"].abs().double().mean() * B * T
    assert torch.all((full - parts).abs() <= 1e-12 * scale), (full - parts).abs().max()
    assert torch.isfinite(full).all() and full.abs().max() > 0


def test_recorded_rollout_invariants_at_full_size(dev):
    B = 1 << 18
    pi0 = _pi0(B, dev, seed=5)
    out = eng.rollout(pi0, THETA, SHIFT, SCALE, T, reward="none", seed=7, outputs=("states", "actions"))
    S, P = out["states"], out["actions"]
    assert S.shape == (T + 1, B, D) and P.shape == (T, B, D, D)
    worst_row, worst_step, worst_mass, pmin = 0.0, 0.0, 0.0, 1.0
    for t in range(T):                                            # one step at a time: 236 MB of actions each
        Pt, St = P[t].double(), S[t].double()
        worst_row = max(worst_row, float((Pt.sum(-1) - 1.0).abs().max()))
        nxt = torch.einsum("bi,bij->bj", St, Pt)
        worst_step = max(worst_step, float((nxt - S[t + 1].double()).abs().max()))
        worst_mass = max(worst_mass, float((S[t + 1].double().sum(-1) - St.sum(-1)).abs().max()))
        pmin = min(pmin, float(Pt.min()))
    assert pmin > 0.0
    assert worst_row <= 1e-6, worst_row              # float P = y * (1/s): 15 roundings of 6e-8
    assert worst_step <= 3e-7, worst_step            # states are stored in float
    assert worst_mass <= 3e-7, worst_mass


def test_td_errors_telescope_at_full_size(dev):
    B = 1 << 18
    pi0, w = _pi0(B, dev, seed=9), _w(dev)
    out = eng.rollout(pi0, THETA, SHIFT, SCALE, T, w=w, gamma=1.0, seed=11,
                      outputs=("rewards", "deltas", "pi_final"), want_acc=True)
    v0, vT = _value(pi0, w), _value(out["pi_final"], w)
    lhs = out["deltas"].double().sum(0)
    rhs = out["rewards"].double().sum(0) + vT - v0
    # deltas / rewards are stored in float (6e-8 each, 16 terms), pi_final in float (V is quadratic in it)
    assert torch.all((lhs - rhs).abs() <= 3e-6 * (1.0 + v0.abs())), float((lhs - rhs).abs().max())
    acc = out["acc"]
    np.testing.assert_allclose(float(acc[-1]), float(out["rewards"].double().sum()), rtol=1e-6)
    # bias feature of the critic gradient = sum of all TD errors
    np.testing.assert_allclose(float(acc[O.num_features(D)]), float(out["deltas"].double().sum()), rtol=1e-5, atol=1e-3)


def test_config2_shape_4096_populations_vs_oracle(dev):
    """BASELINE config 2's shape -- AC_IRL defaults (theta 8.64, shift 0, alpha_scale 1e4, ac_irl.py:33), 4096
    populations x 15 steps -- on injected Gamma variates against the float64 oracle at north_star's 1e-5."""
    B, T2, th, sh, sc = 4096, 15, 8.64, 0.0, 1e4
    rng = np.random.RandomState(2)
    pi0 = np.float32(rng.dirichlet(np.ones(D), size=B))
    w = rng.rand(O.num_features(D))
    y = np.zeros((T2, B, D, D), np.float32)
    pi = pi0.astype(np.float64)
    for t in range(T2):
        alpha, _ = O.policy_alpha(pi, th, sh)
        y[t] = np.float32(rng.gamma(alpha * sc))
        pi = O.mean_field_step(O.normalise_gamma(y[t].astype(np.float64)), pi)
    ref = O.rollout_frozen(pi0.astype(np.float64), th, sh, sc, y.astype(np.float64), w=w, gamma=0.95, discount="cumulative")
    out = eng.rollout(torch.as_tensor(pi0, device=dev), th, sh, sc, T2, w=torch.as_tensor(w, device=dev), gamma=0.95,
                      discount="cumulative", noise_y=torch.as_tensor(y, device=dev),
                      outputs=("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads"), want_acc=True)
    N = lambda t: t.double().cpu().numpy()
    for k in ("states", "actions", "alpha"):
        np.testing.assert_allclose(N(out[k]), ref[k], rtol=1e-5, atol=1e-30, err_msg=k)
    np.testing.assert_allclose(N(out["alpha_deriv"]), ref["alpha_deriv"], rtol=1e-5, atol=3e-8)   # float32 rounding of x
    np.testing.assert_allclose(N(out["grads"]), ref["grads"], rtol=1e-5, atol=2e-6 * D)           # mixed-sign sum of d^2 terms
    v = np.abs(O.features(ref["states"]) @ w)
    scale = np.abs(ref["rewards"]) + v[1:] + v[:-1]
    assert np.all(np.abs(N(out["rewards"]) - ref["rewards"]) <= 1e-6 * np.maximum(scale, 1e-3))
    assert np.all(np.abs(N(out["deltas"]) - ref["deltas"]) <= 1e-6 * scale)
    acc = N(out["acc"])
    F = O.num_features(D)
    np.testing.assert_allclose(acc[0], ref["G_theta"], rtol=1e-5)
    wscale = np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    assert np.all(np.abs(acc[1:1 + F] - ref["G_w"]) <= 1e-6 * wscale + 1e-12)
    np.testing.assert_allclose(acc[1 + F], ref["R"], rtol=1e-5)


def test_config3_sample_of_2_16_populations_vs_oracle(dev):
    """A 2^16-population launch of config 3 (in-kernel Philox, 16 steps, recorded): 2048 randomly chosen populations are
    re-evaluated by the float64 oracle from the RECORDED (pi_t, P_t) -- alpha, alpha', pi', reward, TD error, gradient
    -- at 1e-5.  (The record's float32 rows sum to 1 only to 1e-7, which bounds how close r and delta can be.)"""
    B = 1 << 16
    w = _w(dev)
    out = eng.rollout(_pi0(B, dev, seed=5), THETA, SHIFT, SCALE, T, w=w, seed=4321,
                      outputs=("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads"), want_acc=True)
    idx = torch.as_tensor(np.random.RandomState(1).choice(B, 2048, replace=False), device=dev)
    S = out["states"][:, idx].double().cpu().numpy()
    P = out["actions"][:, idx].double().cpu().numpy()
    alpha, deriv = O.policy_alpha(S[:-1], THETA, SHIFT)
    np.testing.assert_allclose(out["alpha"][:, idx].double().cpu().numpy(), alpha, rtol=1e-5)
    np.testing.assert_allclose(out["alpha_deriv"][:, idx].double().cpu().numpy(), deriv, rtol=1e-5, atol=3e-8)
    np.testing.assert_allclose(np.einsum("tbi,tbij->tbj", S[:-1], P), S[1:], rtol=1e-5, atol=1e-9)
    g = O.log_policy_gradient(alpha, deriv, P)
    np.testing.assert_allclose(out["grads"][:, idx].double().cpu().numpy(), g, rtol=1e-5, atol=2e-6 * D)
    r = O.reward_ac2(P, S[:-1])
    wn = w.cpu().numpy()
    v = O.features(S) @ wn
    delta = r + v[1:] - v[:-1]
    scale = np.abs(r) + np.abs(v[1:]) + np.abs(v[:-1])
    assert np.all(np.abs(out["rewards"][:, idx].double().cpu().numpy() - r) <= 1e-6 * np.maximum(scale, 1e-3))
    assert np.all(np.abs(out["deltas"][:, idx].double().cpu().numpy() - delta) <= 1e-6 * scale)
