"""GPU parity of the rollout / actor-critic kernels (through the C ABI) against
the oracle and the reference-generated golden fixtures.

Tolerances (stated here, used below)
  float64 streams : rtol 1e-10 everywhere (libm-level differences only)
  float32 streams : pi', P, alpha, g      rtol 1e-5  (north_star's tolerance; achieved maxima over 2^16 populations:
                                          pi' 6e-8, P 1.2e-7, alpha 1.9e-6 -- profiles/r2_parity_maxerr.md)
                    alpha', g             quantities that cross zero (alpha' = x sigma(theta x), x = pi_j - pi_i - shift;
                                          g sums d^2 terms of mixed sign): rtol 1e-5 plus an absolute floor from the
                                          float32 rounding of x (|err(x)| <= 3e-8), stated at each use
                    r, delta, sums        |err| <= 1e-6 * operand scale; the kernels accumulate them in float64, so
                                          the error is the fp32 rounding of P / alpha only
                    G_theta               rtol 1e-5 of the sum plus 1e-7 of sum |delta g| (cancellation)
"""
import numpy as np
import pytest
import torch

from oracle import mfg_oracle as O

pytestmark = pytest.mark.gpu

eng = pytest.importorskip("discrete_mean_field_game_b200.engine")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    eng.require_cuda()
    return torch.device("cuda:0")


def T_(x, dev, dtype):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=dev)


def N_(t):
    return t.detach().cpu().numpy().astype(np.float64)


RT = {torch.float64: 1e-10, torch.float32: 1e-5}
ALL_OUT = ("states", "actions", "alpha", "alpha_deriv", "rewards", "deltas", "grads", "pi_final")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("variant", ["fast", "generic"])
def test_kat_d4(dev, kat, dtype, variant):
    """test2.py:6,16,105-121 (d=4, theta=10, shift=0.4) through both kernel variants."""
    th, sh, sc = float(kat["d4_theta"]), float(kat["d4_shift"]), float(kat["d4_alpha_scale"])
    y = T_(kat["d4_y"][None, None], dev, dtype)
    out = eng.rollout(T_(kat["d4_pi"][None], dev, dtype), th, sh, sc, 1, noise_y=y,
                      outputs=("states", "actions", "alpha", "alpha_deriv", "rewards", "grads", "pi_final"),
                      variant=variant)
    rt = RT[dtype]
    # float32 cannot represent the float64 y exactly: compare against the oracle fed the rounded y
    y_used = N_(y)[0, 0]
    o = O.transition(N_(T_(kat["d4_pi"], dev, dtype)), th, sh, sc, y_used)
    np.testing.assert_allclose(N_(out["alpha"])[0, 0], o["alpha"], rtol=rt)
    np.testing.assert_allclose(N_(out["alpha_deriv"])[0, 0], o["alpha_deriv"], rtol=rt)
    np.testing.assert_allclose(N_(out["actions"])[0, 0], o["P"], rtol=rt)
    np.testing.assert_allclose(N_(out["states"])[1, 0], o["pi_next"], rtol=rt)
    np.testing.assert_allclose(N_(out["pi_final"])[0], o["pi_next"], rtol=rt)
    np.testing.assert_allclose(N_(out["rewards"])[0, 0], o["reward"], rtol=10 * rt)
    np.testing.assert_allclose(N_(out["grads"])[0, 0], o["grad"], rtol=rt)
    if dtype == torch.float64:       # and directly against the reference's numbers
        np.testing.assert_allclose(N_(out["alpha"])[0, 0], kat["d4_alpha"], rtol=1e-10)
        np.testing.assert_allclose(N_(out["alpha_deriv"])[0, 0], kat["d4_alpha_deriv"], rtol=1e-10)
        np.testing.assert_allclose(N_(out["actions"])[0, 0], kat["d4_P"], rtol=1e-12)
        np.testing.assert_allclose(N_(out["states"])[1, 0], kat["d4_pi_next"], rtol=1e-12)
        np.testing.assert_allclose(N_(out["rewards"])[0, 0], kat["d4_reward"], rtol=1e-10)
        np.testing.assert_allclose(N_(out["grads"])[0, 0], kat["d4_grad_vectorized"], rtol=1e-10)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_kat_d3_rewards(dev, kat, dtype):
    """calc_reward on the fixed 3x3 input (test2.py:46-56) -- feed P as 'Gamma variates' whose rows
    the kernel normalises, so use the row-stochastic version and compare with the oracle."""
    P = kat["P3_rowstochastic"]
    pi = kat["pi3"]
    for kind, fn in (("ac2", O.reward_ac2), ("synthetic", O.reward_synthetic)):
        out = eng.rollout(T_(pi[None], dev, dtype), 8.0, 0.1, 100.0, 1, noise_y=T_(P[None, None], dev, dtype),
                          reward=kind, outputs=("rewards", "actions"), variant="generic")
        Pn = N_(T_(P, dev, dtype))
        Pn = Pn / Pn.sum(-1, keepdims=True)
        np.testing.assert_allclose(N_(out["rewards"])[0, 0], fn(Pn, N_(T_(pi, dev, dtype))), rtol=20 * RT[dtype])
    assert np.isclose(O.reward_synthetic(P, pi), kat["reward3_synthetic"], rtol=1e-14)


def _trace_frozen_inputs(trace, n_pop=3):
    """B populations = the first episodes of the reference trace, time-major."""
    E = min(n_pop, int(trace["episodes"]))
    y = np.ascontiguousarray(np.transpose(trace["y"][:E], (1, 0, 2, 3)))     # [T,B,d,d]
    pi0 = trace["mat_pi0"][trace["start_rows"][:E]]
    return pi0, y


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("variant", ["fast", "generic"])
@pytest.mark.parametrize("discount", ["step", "cumulative"])
def test_frozen_rollout_d15_vs_oracle(dev, trace, dtype, variant, discount):
    """Config 1 inputs (the reference's own Gamma draws), frozen (theta0, w0): every per-step
    quantity and the reduced sums against the float64 oracle."""
    pi0, y = _trace_frozen_inputs(trace)
    th, sh, sc = float(trace["theta0"]), float(trace["shift"]), float(trace["alpha_scale"])
    gamma = 0.9 if discount == "cumulative" else 1.0
    w = torch.as_tensor(trace["w0"], dtype=torch.float64, device=dev)
    out = eng.rollout(T_(pi0, dev, dtype), th, sh, sc, y.shape[0], w=w, gamma=gamma, discount=discount,
                      noise_y=T_(y, dev, dtype), outputs=ALL_OUT, want_acc=True, variant=variant)
    ref = O.rollout_frozen(N_(T_(pi0, dev, dtype)), th, sh, sc, N_(T_(y, dev, dtype)), w=trace["w0"],
                           gamma=gamma, discount=discount)
    rt = RT[dtype]
    for k in ("states", "actions", "alpha", "alpha_deriv", "grads"):
        np.testing.assert_allclose(N_(out[k]), ref[k], rtol=rt, err_msg=k)
    np.testing.assert_allclose(N_(out["pi_final"]), ref["states"][-1], rtol=rt)
    # r and delta: differences of nearly equal terms -> operand-scaled tolerance (hard part 2)
    v = np.abs(O.features(ref["states"]) @ trace["w0"])
    scale = np.abs(ref["rewards"]) + v[1:] + v[:-1]
    tol = 1e-10 if dtype == torch.float64 else 1e-6
    assert np.all(np.abs(N_(out["rewards"]) - ref["rewards"]) <= tol * np.maximum(scale, 1e-3))
    assert np.all(np.abs(N_(out["deltas"]) - ref["deltas"]) <= tol * scale)
    acc = N_(out["acc"])
    F = O.num_features(15)
    gscale = np.sum(np.abs(ref["deltas"] * ref["grads"]))
    assert abs(acc[0] - ref["G_theta"]) <= 10 * tol * gscale
    wscale = np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    assert np.all(np.abs(acc[1:1 + F] - ref["G_w"]) <= 10 * tol * wscale)
    assert abs(acc[1 + F] - ref["R"]) <= 10 * tol * np.sum(np.abs(ref["rewards"]))
    # and the very first transition against the reference's own numbers (nothing updated yet)
    if dtype == torch.float64 and discount == "step":
        np.testing.assert_allclose(N_(out["actions"])[0, 0], trace["P"][0, 0], rtol=1e-12)
        np.testing.assert_allclose(N_(out["deltas"])[0, 0], trace["delta"][0, 0], rtol=1e-9)
        np.testing.assert_allclose(N_(out["grads"])[0, 0], trace["grad"][0, 0], rtol=1e-10)
        np.testing.assert_allclose(N_(out["rewards"])[0, 0], trace["reward"][0, 0], rtol=1e-9)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_learners_replay_reference_train(dev, trace, dtype):
    """Config 1: mfg_ac2.actor_critic.train replayed on the GPU with the reference's own start rows and
    Gamma draws -- per-step theta and the final critic weights against the reference's numbers."""
    E, T, d = int(trace["episodes"]), int(trace["T"]), int(trace["d"])
    theta = torch.tensor([float(trace["theta0"])], dtype=torch.float64, device=dev)
    w = torch.as_tensor(trace["w0"][None], dtype=torch.float64, device=dev).contiguous()
    res = eng.learners(theta, w, T_(trace["mat_pi0"], dev, dtype), E, T, shift=float(trace["shift"]),
                       alpha_scale=float(trace["alpha_scale"]), episode0=0, gamma=1.0, lr_critic=0.1,
                       lr_actor=0.1, constant=False, reward="ac2", discount="step",
                       start_rows=torch.as_tensor(trace["start_rows"][None], dtype=torch.int32, device=dev),
                       noise_y=T_(trace["y"][None], dev, dtype), trace=True)
    if dtype == torch.float64:
        np.testing.assert_allclose(N_(res["theta_trace"])[0], trace["theta_after"], rtol=1e-11)
        np.testing.assert_allclose(N_(res["delta_trace"])[0], trace["delta"], rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(N_(w)[0], trace["w_final"], rtol=1e-11)
        np.testing.assert_allclose(N_(theta)[0], float(trace["theta_final"]), rtol=1e-12)
        np.testing.assert_allclose(N_(res["total_reward"])[0], trace["reward"].sum(1), rtol=1e-9)
        np.testing.assert_allclose(N_(res["pi_final"])[0], trace["pi_next"][-1, -1], rtol=1e-11)
    else:
        # float32 streams: the update increments are what the 1e-5 target is about
        dth_ref = np.diff(np.concatenate([[float(trace["theta0"])], trace["theta_after"].ravel()]))
        dth = np.diff(np.concatenate([[float(trace["theta0"])], N_(res["theta_trace"]).ravel()]))
        assert np.max(np.abs(dth - dth_ref)) <= 2e-5 * np.max(np.abs(dth_ref))
        np.testing.assert_allclose(N_(theta)[0], float(trace["theta_final"]), rtol=1e-6)
        np.testing.assert_allclose(N_(w)[0], trace["w_final"], rtol=1e-5, atol=1e-7)


def test_learners_are_independent_and_match_serial_oracle(dev, trace):
    """Several learners with different (theta0, shift) in one launch == each one run alone by the oracle
    (the mfg_synthetic.py:902-925 sweep pattern)."""
    E, T, d = 2, 15, 15
    rng = np.random.RandomState(5)
    L = 5
    theta0 = np.array([8.86349, 6.5, 7.4, 9.9, 8.0])
    shifts = np.array([0.16, 0.0, 0.3, 0.16, 0.5])
    w0 = rng.rand(L, O.num_features(d))
    start = rng.randint(0, trace["mat_pi0"].shape[0], size=(L, E)).astype(np.int32)
    y = rng.gamma(shape=50.0, size=(L, E, T, d, d))
    theta = torch.as_tensor(theta0, dtype=torch.float64, device=dev).clone()
    w = torch.as_tensor(w0, dtype=torch.float64, device=dev).clone()
    eng.learners(theta, w, T_(trace["mat_pi0"], dev, torch.float64), E, T,
                 shift=torch.as_tensor(shifts, dtype=torch.float64, device=dev), alpha_scale=12000.0,
                 episode0=1, gamma=0.95, lr_critic=0.1, lr_actor=0.001, constant=False, reward="synthetic",
                 discount="cumulative", start_rows=torch.as_tensor(start, device=dev),
                 noise_y=T_(y, dev, torch.float64))
    for l in range(L):
        th, ww, _ = O.train_serial(trace["mat_pi0"], theta0[l], w0[l], shifts[l], 12000.0, E, gamma=0.95,
                                   lr_critic=0.1, lr_actor=0.001, flavour="ac_irl", reward="synthetic",
                                   noise=O.InjectedNoise(start[l], y[l]), num_steps=T)
        np.testing.assert_allclose(N_(theta)[l], th, rtol=1e-11)
        np.testing.assert_allclose(N_(w)[l], ww, rtol=1e-10)


@pytest.mark.parametrize("d,layout", [(15, "groups"), (16, "groups"), (21, "groups"), (15, "cta"), (16, "cta"), (21, "cta")])
def test_learners_v2_float_match_serial_oracle(dev, d, layout):
    """learners_v2_kernel (float streams; 16-lane groups at d = 15 / 16, 32-lane groups at the reference's default
    d = 21) and learner_cta_kernel (a CTA per learner, d = 15 / 16 / 21): several learners with their own
    (theta0, shift), injected float32 Gamma variates and start rows, both rewards / discount flavours, against the
    float64 serial oracle of mfg_ac2.train / AC_IRL.train."""
    E, T, L = 2, 15, 19                                   # 19 learners: not a multiple of the learners per CTA
    rng = np.random.RandomState(100 + d)
    mat = np.float32(rng.dirichlet(np.ones(d), size=9))
    theta0 = rng.uniform(6.0, 10.0, size=L)
    shifts = rng.uniform(0.0, 0.4, size=L)
    w0 = rng.rand(L, O.num_features(d))
    start = rng.randint(0, mat.shape[0], size=(L, E)).astype(np.int32)
    y = np.float32(rng.gamma(shape=50.0, size=(L, E, T, d, d)))
    y[3, 1, 4, 2, 5] = 0.0                                # an exact zero variate (mfg_ac2.py:244)
    for reward, discount, flavour, gamma in (("ac2", "step", "mfg_ac2", 1.0), ("synthetic", "cumulative", "ac_irl", 0.95)):
        theta = torch.as_tensor(theta0, dtype=torch.float64, device=dev).clone()
        w = torch.as_tensor(w0, dtype=torch.float64, device=dev).clone()
        ep0 = 0 if flavour == "mfg_ac2" else 1
        res = eng.learners(theta, w, T_(mat, dev, torch.float32), E, T,
                           shift=torch.as_tensor(shifts, dtype=torch.float64, device=dev), alpha_scale=12000.0,
                           episode0=ep0, gamma=gamma, lr_critic=0.1, lr_actor=0.01, constant=False, reward=reward,
                           discount=discount, start_rows=torch.as_tensor(start, device=dev),
                           noise_y=T_(y, dev, torch.float32), layout=layout)
        for l in range(L):
            th, ww, info = O.train_serial(mat.astype(np.float64), theta0[l], w0[l], shifts[l], 12000.0, E, gamma=gamma,
                                          lr_critic=0.1, lr_actor=0.01, flavour=flavour, reward=reward,
                                          noise=O.InjectedNoise(start[l], y[l].astype(np.float64)), num_steps=T)
            np.testing.assert_allclose(N_(theta)[l], th, rtol=2e-6, err_msg="theta of learner %d" % l)
            np.testing.assert_allclose(N_(w)[l], ww, rtol=2e-5, atol=2e-6, err_msg="w of learner %d" % l)


@pytest.mark.parametrize("d", [15, 16, 21])
def test_learner_cta_and_group_kernels_agree_on_philox_draws(dev, d):
    """The two layouts of dmfg_ac_learners key their draws identically (seed, learner, step, row, pair): with the
    in-kernel sampler a learner follows the same trajectory in both, up to the order of the double sums -- theta and
    delta after every step, episode rewards, final state and weights; AUTO picks the CTA form for a few learners."""
    L, E, T = 5, 3, 15
    rng = np.random.RandomState(d)
    mat = T_(np.float32(rng.dirichlet(np.ones(d), size=11)), dev, torch.float32)
    th0 = rng.uniform(6.0, 10.0, size=L)
    w0 = rng.rand(L, O.num_features(d))
    outs = {}
    for layout in ("groups", "cta", "auto"):
        theta = torch.as_tensor(th0, dtype=torch.float64, device=dev).clone()
        w = torch.as_tensor(w0, dtype=torch.float64, device=dev).clone()
        res = eng.learners(theta, w, mat, E, T, shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.01, seed=77,
                           trace=True, layout=layout, learner_offset=3)
        outs[layout] = (N_(theta), N_(w), N_(res["theta_trace"]), N_(res["delta_trace"]), N_(res["total_reward"]),
                        N_(res["pi_final"]))
    g, c, a = outs["groups"], outs["cta"], outs["auto"]
    for x, y in zip(c, a):
        np.testing.assert_array_equal(x, y)                     # AUTO == CTA at L = 5
    np.testing.assert_allclose(c[0], g[0], rtol=1e-9)
    np.testing.assert_allclose(c[1], g[1], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(c[2], g[2], rtol=1e-9)
    np.testing.assert_allclose(c[3], g[3], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(c[4], g[4], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(c[5], g[5], rtol=1e-6)


def test_generate_trajectory_d15(dev, traj15):
    out = eng.rollout(T_(traj15["pi0"][None], dev, torch.float64), float(traj15["theta"]),
                      float(traj15["shift"]), float(traj15["alpha_scale"]), 15, reward="none",
                      noise_y=T_(traj15["y"][:, None], dev, torch.float64), outputs=("states",))
    np.testing.assert_allclose(N_(out["states"])[:, 0], traj15["trajectory"], rtol=1e-11)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_forward_d47_generic(dev, fwd47, dtype):
    """test2.py:203-224 start state (d=47) through the generic kernel + TD pass."""
    d = 47
    rng = np.random.RandomState(3)
    w0 = rng.rand(O.num_features(d))
    out = eng.rollout(T_(fwd47["states"][:1], dev, dtype), float(fwd47["theta"]), float(fwd47["shift"]),
                      float(fwd47["alpha_scale"]), 3, w=torch.as_tensor(w0, dtype=torch.float64, device=dev),
                      noise_y=T_(fwd47["y"][:, None], dev, dtype), outputs=ALL_OUT, want_acc=True)
    rt = RT[dtype]
    if dtype == torch.float64:
        np.testing.assert_allclose(N_(out["states"])[:, 0], fwd47["states"], rtol=1e-11)
        np.testing.assert_allclose(N_(out["actions"])[:, 0], fwd47["actions"], rtol=1e-11)
        np.testing.assert_allclose(N_(out["rewards"])[:, 0], fwd47["reward"], rtol=1e-9)
        np.testing.assert_allclose(N_(out["grads"])[:, 0], fwd47["grad"], rtol=1e-10)
    ref = O.rollout_frozen(N_(T_(fwd47["states"][:1], dev, dtype)), float(fwd47["theta"]), float(fwd47["shift"]),
                           float(fwd47["alpha_scale"]), N_(T_(fwd47["y"][:, None], dev, dtype)), w=w0)
    np.testing.assert_allclose(N_(out["states"]), ref["states"], rtol=rt)
    np.testing.assert_allclose(N_(out["grads"]), ref["grads"], rtol=rt)
    tol = 1e-10 if dtype == torch.float64 else 1e-6
    v = np.abs(O.features(ref["states"]) @ w0)
    assert np.all(np.abs(N_(out["deltas"]) - ref["deltas"]) <= tol * (np.abs(ref["rewards"]) + v[1:] + v[:-1]))
    F = O.num_features(d)
    acc = N_(out["acc"])
    np.testing.assert_allclose(acc[1:1 + F], ref["G_w"], rtol=1e3 * tol, atol=1e-12)
    np.testing.assert_allclose(acc[0], ref["G_theta"], rtol=1e3 * tol)
    np.testing.assert_allclose(acc[1 + F], ref["R"], rtol=1e3 * tol)


def test_td_accumulate_matches_fused(dev, trace):
    """dmfg_td_accumulate on a recorded batch == the fused accumulators of dmfg_rollout (and external
    rewards replace the closed form, as the reward net does in ac_irl.py:683)."""
    pi0, y = _trace_frozen_inputs(trace)
    w = torch.as_tensor(trace["w0"], dtype=torch.float64, device=dev)
    a = eng.rollout(T_(pi0, dev, torch.float64), 8.86349, 0.16, 12000.0, 15, w=w,
                    noise_y=T_(y, dev, torch.float64), outputs=("states", "rewards", "grads", "deltas"),
                    want_acc=True)
    b = eng.td_accumulate(a["states"], a["rewards"], a["grads"], w)
    np.testing.assert_allclose(N_(b["deltas"]), N_(a["deltas"]), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(N_(b["acc"]), N_(a["acc"]), rtol=1e-9, atol=1e-14)
    ext = torch.rand_like(a["rewards"]) - 0.5
    c = eng.rollout(T_(pi0, dev, torch.float64), 8.86349, 0.16, 12000.0, 15, w=w, rewards_in=ext,
                    noise_y=T_(y, dev, torch.float64), outputs=("deltas",), want_acc=True)
    e = eng.td_accumulate(a["states"], ext, a["grads"], w)
    np.testing.assert_allclose(N_(e["deltas"]), N_(c["deltas"]), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(N_(e["acc"]), N_(c["acc"]), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("d", [4, 15, 16])
def test_philox_fast_equals_generic(dev, d):
    """Draws are keyed by (seed, population id, step, row, pair): both kernel variants, any launch
    geometry and any split of the batch give the same trajectories."""
    B, T = 37, 5
    rng = np.random.RandomState(d)
    pi0 = T_(rng.dirichlet(np.ones(d), size=B), dev, torch.float32)
    kw = dict(seed=1234, outputs=("states", "actions", "rewards", "grads"))
    f = eng.rollout(pi0, 8.64, 0.0, 1e4, T, variant="fast", **kw)
    g = eng.rollout(pi0, 8.64, 0.0, 1e4, T, variant="generic", **kw)
    for k in kw["outputs"]:
        # g is a sum of d^2 terms of mixed sign (|terms| ~ 1): the two kernels sum it differently (per-element ln P
        # vs. the unnormalised form sum alpha' lg2 y - lg2 s sum alpha'), so its bound is absolute, 2e-6 ~ 20 float ulps of a term
        atol = 2e-6 if k == "grads" else 1e-7
        np.testing.assert_allclose(N_(f[k]), N_(g[k]), rtol=2e-5, atol=atol, err_msg=k)
    # shard invariance: the second half alone, addressed by its global population ids
    h = eng.rollout(pi0[20:].contiguous(), 8.64, 0.0, 1e4, T, variant="fast", pop_offset=20, **kw)
    assert torch.equal(h["states"], f["states"][:, 20:])
    assert torch.equal(h["actions"], f["actions"][:, 20:])
    # a different seed gives different draws
    k2 = eng.rollout(pi0, 8.64, 0.0, 1e4, T, variant="fast", seed=99, outputs=("actions",))
    assert not torch.equal(k2["actions"], f["actions"])


@pytest.mark.parametrize("d", [15, 21, 32, 47, 64, 255, 256])
def test_philox_rollout_invariants(dev, d):
    """test2.py:26,32 and test_acirl.py:43-47: rows of P sum to 1, mass is conserved,
    state_{t+1} = action_t^T state_t -- at a size the oracle never sees."""
    B, T = 4096 if d < 64 else 256, 16
    rng = np.random.RandomState(0)
    pi0 = T_(rng.dirichlet(np.ones(d), size=B), dev, torch.float32)
    out = eng.rollout(pi0, 8.86349, 0.16, 12000.0, T, seed=7, outputs=("states", "actions"))
    P, S = out["actions"].double(), out["states"].double()
    assert torch.all(P > 0)
    assert torch.allclose(P.sum(-1), torch.ones_like(P.sum(-1)), atol=5e-7)
    assert torch.allclose(S.sum(-1), torch.ones_like(S.sum(-1)), atol=2e-6)
    nxt = torch.einsum("tbi,tbij->tbj", S[:-1], P)
    assert torch.allclose(nxt, S[1:], atol=3e-7)


def test_empty_and_degenerate_sizes(dev):
    z = torch.zeros((0, 15), dtype=torch.float32, device=dev)
    out = eng.rollout(z, 8.0, 0.1, 1e4, 16, outputs=("states", "actions"))
    assert out["states"].shape == (17, 0, 15) and out["actions"].shape == (16, 0, 15, 15)
    one = torch.full((1, 15), 1 / 15, dtype=torch.float32, device=dev)
    out = eng.rollout(one, 8.0, 0.1, 1e4, 0, outputs=("states", "pi_final"))
    assert torch.equal(out["states"][0], one) and torch.equal(out["pi_final"], one)
    w = torch.zeros(136, dtype=torch.float64, device=dev)
    out = eng.rollout(z, 8.0, 0.1, 1e4, 4, w=w, want_acc=True, outputs=())
    assert torch.count_nonzero(out["acc"]) == 0
    # ragged: B not a multiple of the 16 populations a CTA handles
    for B in (1, 15, 17, 33):
        pi0 = torch.softmax(torch.randn(B, 15, device=dev), -1)
        out = eng.rollout(pi0, 8.0, 0.1, 1e4, 3, w=w + 0.5, want_acc=True, outputs=("states", "deltas"))
        assert torch.isfinite(out["states"]).all() and torch.isfinite(out["acc"]).all()


def test_zero_gamma_variate_is_replaced(dev):
    """y == 0 -> 1e-20 (mfg_ac2.py:244) and ln P stays finite."""
    d = 4
    y = torch.ones((1, 1, d, d), dtype=torch.float64, device=dev)
    y[0, 0, 1, 2] = 0.0
    pi0 = torch.tensor([[0.4, 0.3, 0.2, 0.1]], dtype=torch.float64, device=dev)
    out = eng.rollout(pi0, 10.0, 0.4, 12000.0, 1, noise_y=y, outputs=("actions", "grads"))
    ref = O.transition(N_(pi0)[0], 10.0, 0.4, 12000.0, N_(y)[0, 0])
    np.testing.assert_allclose(N_(out["actions"])[0, 0], ref["P"], rtol=1e-12)
    np.testing.assert_allclose(N_(out["grads"])[0, 0], ref["grad"], rtol=1e-10)
    assert N_(out["actions"])[0, 0, 1, 2] > 0


def test_per_episode_update_on_device(dev, trace):
    """dmfg_ac_apply_update: theta += lr_a/B * acc[0], w += lr_c/B * acc[1:]."""
    pi0, y = _trace_frozen_inputs(trace)
    w = torch.as_tensor(trace["w0"], dtype=torch.float64, device=dev).clone()
    theta = torch.tensor([8.86349], dtype=torch.float64, device=dev)
    out = eng.rollout(T_(pi0, dev, torch.float64), 0.0, 0.16, 12000.0, 15, w=w, theta_dev=theta,
                      noise_y=T_(y, dev, torch.float64), outputs=(), want_acc=True)
    acc = N_(out["acc"])
    ref = O.rollout_frozen(pi0, 8.86349, 0.16, 12000.0, y, w=trace["w0"])
    np.testing.assert_allclose(acc[0], ref["G_theta"], rtol=1e-9)
    eng.apply_update(15, theta, w, out["acc"], 0.1, 0.01, 1.0 / 3)
    np.testing.assert_allclose(N_(theta)[0], 8.86349 + 0.01 / 3 * acc[0], rtol=1e-14)
    np.testing.assert_allclose(N_(w), trace["w0"] + 0.1 / 3 * acc[1:137], rtol=1e-14)


# ----------------------------------------------------------------------------- throughput kernel (v2)
@pytest.mark.parametrize("discount", ["step", "cumulative"])
@pytest.mark.parametrize("reward", ["ac2", "synthetic"])
def test_v2_frozen_rollout_d15_vs_oracle(dev, trace, discount, reward):
    """The kernel bench.py times (rollout_v2_kernel, float streams) on the reference's own Gamma draws:
    same checks and the same tolerances as the float32 rows of test_frozen_rollout_d15_vs_oracle."""
    pi0, y = _trace_frozen_inputs(trace)
    th, sh, sc = float(trace["theta0"]), float(trace["shift"]), float(trace["alpha_scale"])
    gamma = 0.9 if discount == "cumulative" else 1.0
    dtype = torch.float32
    w = torch.as_tensor(trace["w0"], dtype=torch.float64, device=dev)
    out = eng.rollout(T_(pi0, dev, dtype), th, sh, sc, y.shape[0], w=w, gamma=gamma, discount=discount,
                      reward=reward, noise_y=T_(y, dev, dtype), outputs=ALL_OUT, want_acc=True, variant="v2")
    ref = O.rollout_frozen(N_(T_(pi0, dev, dtype)), th, sh, sc, N_(T_(y, dev, dtype)), w=trace["w0"],
                           gamma=gamma, discount=discount, reward=reward)
    rt = RT[dtype]
    for k in ("states", "actions", "alpha", "alpha_deriv", "grads"):
        np.testing.assert_allclose(N_(out[k]), ref[k], rtol=rt, err_msg=k)
    np.testing.assert_allclose(N_(out["pi_final"]), ref["states"][-1], rtol=rt)
    v = np.abs(O.features(ref["states"]) @ trace["w0"])
    scale = np.abs(ref["rewards"]) + v[1:] + v[:-1]
    tol = 1e-6
    assert np.all(np.abs(N_(out["rewards"]) - ref["rewards"]) <= tol * np.maximum(scale, 1e-3))
    assert np.all(np.abs(N_(out["deltas"]) - ref["deltas"]) <= tol * scale)
    acc = N_(out["acc"])
    F = O.num_features(15)
    assert abs(acc[0] - ref["G_theta"]) <= 10 * tol * np.sum(np.abs(ref["deltas"] * ref["grads"]))
    wscale = np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    assert np.all(np.abs(acc[1:1 + F] - ref["G_w"]) <= 10 * tol * wscale)
    assert abs(acc[1 + F] - ref["R"]) <= 10 * tol * np.sum(np.abs(ref["rewards"]))


@pytest.mark.parametrize("d,T", [(15, 6), (16, 6), (20, 6), (21, 6), (21, 7), (21, 9), (32, 5), (32, 8)])
def test_v2_random_batch_vs_oracle(dev, d, T):
    """d = 15 / 16 (16-lane groups) and d = 21 / 32 (32-lane groups), ~200 populations (not a multiple of the
    16- / 8-population tile), step counts that end at every position of the Gram flush group, random Gamma variates
    including shapes below 1 and an exact zero."""
    rng = np.random.RandomState(d)
    B = 200 if d <= 16 else 203
    pi0 = np.float32(rng.dirichlet(np.ones(d) * 0.7, size=B))
    F = O.num_features(d)
    w = rng.rand(F)
    # draw the variates along the oracle's own trajectory, rounded to float32
    y = np.zeros((T, B, d, d), np.float32)
    pi = pi0.astype(np.float64)
    for t in range(T):
        alpha, _ = O.policy_alpha(pi, 8.64, 0.05)
        y[t] = np.float32(rng.gamma(alpha * 1e4))
        if t == 2:
            y[t, 5, 3, 7] = 0.0
        pi = O.mean_field_step(O.normalise_gamma(y[t].astype(np.float64)), pi)
    ref = O.rollout_frozen(pi0.astype(np.float64), 8.64, 0.05, 1e4, y.astype(np.float64), w=w)
    out = eng.rollout(T_(pi0, dev, torch.float32), 8.64, 0.05, 1e4, T, w=T_(w, dev, torch.float64),
                      noise_y=T_(y, dev, torch.float32), outputs=ALL_OUT, want_acc=True, variant="v2")
    for k in ("states", "alpha"):
        np.testing.assert_allclose(N_(out[k]), ref[k], rtol=1e-5, err_msg=k)
    # alpha' = x sigma(theta x) crosses zero with x = pi_j - pi_i - shift: float32 inputs bound |err(x)| by ~1e-8
    np.testing.assert_allclose(N_(out["alpha_deriv"]), ref["alpha_deriv"], rtol=1e-5, atol=3e-8)
    np.testing.assert_allclose(N_(out["actions"]), ref["actions"], rtol=1e-5, atol=1e-30)
    # g sums d^2 terms of mixed sign: bound relative to the sum of their magnitudes (~ d for these inputs)
    np.testing.assert_allclose(N_(out["grads"]), ref["grads"], rtol=1e-5, atol=max(2e-5, 2e-6 * d))
    v = np.abs(O.features(ref["states"]) @ w)
    scale = np.abs(ref["rewards"]) + v[1:] + v[:-1]
    assert np.all(np.abs(N_(out["deltas"]) - ref["deltas"]) <= 1e-6 * scale)
    assert np.all(np.abs(N_(out["rewards"]) - ref["rewards"]) <= 1e-6 * np.maximum(scale, 1e-3))
    acc = N_(out["acc"])
    assert abs(acc[0] - ref["G_theta"]) <= 1e-5 * abs(ref["G_theta"]) + 1e-7 * np.sum(np.abs(ref["deltas"] * ref["grads"]))
    wscale = np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    assert np.all(np.abs(acc[1:1 + F] - ref["G_w"]) <= 1e-5 * wscale + 1e-12)
    np.testing.assert_allclose(acc[1 + F], ref["R"], rtol=1e-5, atol=1e-9)
    # the TRAIN specialisation (no per-step stream) on the same variates: same sums
    tr = eng.rollout(T_(pi0, dev, torch.float32), 8.64, 0.05, 1e4, T, w=T_(w, dev, torch.float64),
                     noise_y=T_(y, dev, torch.float32), outputs=(), want_acc=True, variant="v2")
    assert np.all(np.abs(N_(tr["acc"])[1:1 + F] - acc[1:1 + F]) <= 1e-9 * wscale + 1e-12)


@pytest.mark.parametrize("d", [15, 16, 20, 21, 32])
def test_v2_philox_draws_match_other_variants(dev, d):
    """Same (seed, population, step, row, pair) -> same Gamma variates in every kernel variant (d = 21 / 32: the
    32-lane v2 kernel against the wide kernel, which is what "generic" runs for float streams)."""
    B, T = 53, 4
    rng = np.random.RandomState(d + 1)
    pi0 = T_(rng.dirichlet(np.ones(d), size=B), dev, torch.float32)
    kw = dict(seed=4321, outputs=("states", "actions", "rewards", "grads"))
    a = eng.rollout(pi0, 8.64, 0.0, 1e4, T, variant="v2", **kw)
    b = eng.rollout(pi0, 8.64, 0.0, 1e4, T, variant="fast" if d <= 16 else "generic", **kw)
    for k in ("states", "actions"):
        np.testing.assert_allclose(N_(a[k]), N_(b[k]), rtol=3e-5, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(N_(a["grads"]), N_(b["grads"]), rtol=1e-4, atol=1e-4)
    h = eng.rollout(pi0[16:].contiguous(), 8.64, 0.0, 1e4, T, variant="v2", pop_offset=16, **kw)
    assert torch.equal(h["actions"], a["actions"][:, 16:])
    auto = eng.rollout(pi0, 8.64, 0.0, 1e4, T, **kw)                       # AUTO picks v2 here
    assert torch.equal(auto["actions"], a["actions"])


@pytest.mark.parametrize("d,B,T", [(32, 37, 5), (64, 203, 3), (48, 16, 16), (80, 5, 17), (128, 3, 1), (32, 1, 1)])
@pytest.mark.parametrize("discount", ["step", "cumulative"])
def test_td_pass_on_fp64_tensor_cores_matches_numpy(dev, d, B, T, discount):
    """Wide states (d a multiple of 16, float streams): V = phi(pi).w and sum delta*phi run as DMMA GEMMs
    (dmfg_td_dmma.cuh).  Sizes with ragged tiles (states not a multiple of 8, transitions not a multiple of 4);
    checked against the float64 formulas of mfg_ac2.py:290-344, 505-514 on the same recorded batch."""
    rng = np.random.RandomState(d + B)
    F = O.num_features(d)
    w = rng.randn(F)
    pi0 = T_(rng.dirichlet(np.ones(d), size=B), dev, torch.float32)
    rec = eng.rollout(pi0, 8.0, 0.1, 1e4, T, reward="ac2", seed=3, outputs=("states", "rewards", "grads"),
                      variant="generic")
    gamma = 0.9
    wd = torch.as_tensor(w, dtype=torch.float64, device=dev)
    td = eng.td_accumulate(rec["states"], rec["rewards"], rec["grads"], wd, gamma=gamma, discount=discount)
    S, r, g = N_(rec["states"]), N_(rec["rewards"]), N_(rec["grads"])
    phi = O.features(S)                                           # [T+1, B, F]
    V = phi @ w
    gf = np.full(T, gamma) if discount == "step" else gamma ** np.arange(T)
    delta = r + gf[:, None] * V[1:] - V[:-1]
    vs = np.abs(V[1:]) + np.abs(V[:-1]) + np.abs(r)
    assert np.all(np.abs(N_(td["deltas"]) - delta) <= 1e-6 * vs + 1e-12)       # deltas are stored in float
    acc = N_(td["acc"])
    G_w = np.einsum("tb,tbf->f", delta, phi[:-1])
    scale = np.einsum("tb,tbf->f", np.abs(delta), np.abs(phi[:-1]))
    assert np.all(np.abs(acc[1:1 + F] - G_w) <= 1e-10 * scale + 1e-15), np.max(np.abs(acc[1:1 + F] - G_w) / (scale + 1e-300))
    np.testing.assert_allclose(acc[0], np.sum(delta * g), rtol=1e-10)
    np.testing.assert_allclose(acc[1 + F], np.sum(r), rtol=1e-12)
    # the fused call (rollout + TD) gives the same sums as the two-step call on its own record
    # (variant "generic": at d = 32 AUTO would pick the 32-lane v2 kernel, which never materialises the record)
    full = eng.rollout(pi0, 8.0, 0.1, 1e4, T, w=wd, gamma=gamma, discount=discount, reward="ac2", seed=3,
                       outputs=("deltas",), want_acc=True, variant="generic")
    np.testing.assert_allclose(N_(full["acc"]), acc, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(N_(full["deltas"]), N_(td["deltas"]), rtol=0, atol=0)


@pytest.mark.parametrize("d,B,T", [(15, 1, 1), (15, 133, 15), (16, 257, 16), (15, 4097, 3)])
@pytest.mark.parametrize("discount", ["step", "cumulative"])
def test_td_pass_of_a_record_at_small_d_matches_numpy(dev, d, B, T, discount):
    """dmfg_td_accumulate on a float32 record at d = 15 / 16 (td_delta_small_kernel: a thread per population, critic
    weights in shared memory; then td_gw_kernel) against the float64 formulas of mfg_ac2.py:290-344, 505-514 / the
    external rewards of ac_irl.py:683-700; ragged batch sizes, both discount conventions."""
    rng = np.random.RandomState(7 * d + B)
    F = O.num_features(d)
    w = rng.randn(F)
    pi0 = T_(rng.dirichlet(np.ones(d), size=B), dev, torch.float32)
    rec = eng.rollout(pi0, 8.0, 0.1, 1e4, T, reward="none", seed=3, outputs=("states", "grads"))
    r_ext = torch.as_tensor(rng.rand(T, B) - 0.5, dtype=torch.float32, device=dev)      # rewards as a reward net would give
    gamma = 0.9
    wd = torch.as_tensor(w, dtype=torch.float64, device=dev)
    td = eng.td_accumulate(rec["states"], r_ext, rec["grads"], wd, gamma=gamma, discount=discount)
    S, r, g = N_(rec["states"]), N_(r_ext), N_(rec["grads"])
    phi = O.features(S)
    V = phi @ w
    gf = np.full(T, gamma) if discount == "step" else gamma ** np.arange(T)
    delta = r + gf[:, None] * V[1:] - V[:-1]
    vs = np.abs(V[1:]) + np.abs(V[:-1]) + np.abs(r)
    assert np.all(np.abs(N_(td["deltas"]) - delta) <= 1e-6 * vs + 1e-12)       # deltas are stored in float
    acc = N_(td["acc"])
    G_w = np.einsum("tb,tbf->f", delta, phi[:-1])
    scale = np.einsum("tb,tbf->f", np.abs(delta), np.abs(phi[:-1]))
    assert np.all(np.abs(acc[1:1 + F] - G_w) <= 1e-10 * scale + 1e-15)
    np.testing.assert_allclose(acc[0], np.sum(delta * g), rtol=1e-10)
    np.testing.assert_allclose(acc[1 + F], np.sum(r), rtol=1e-12)


@pytest.mark.parametrize("d", [21, 100, 200, 256])
def test_wide_kernel_vs_oracle(dev, d):
    """rollout_wide_kernel (float streams, any d; 1 / 2 / 4 column pairs per lane, odd and maximum d) with injected
    Gamma variates against the float64 oracle: every per-step stream and the reduced sums (the TD pass runs on the
    scalar kernels at d = 21 / 100 / 200 and on the FP64 tensor cores at d = 256)."""
    rng = np.random.RandomState(d)
    B, T = 3, 2
    pi0 = np.float32(rng.dirichlet(np.ones(d) * 0.7, size=B))
    F = O.num_features(d)
    w = rng.rand(F)
    y = np.zeros((T, B, d, d), np.float32)
    pi = pi0.astype(np.float64)
    for t in range(T):
        alpha, _ = O.policy_alpha(pi, 8.64, 0.05)
        y[t] = np.float32(rng.gamma(alpha * 1e4))
        pi = O.mean_field_step(O.normalise_gamma(y[t].astype(np.float64)), pi)
    y[1, 2, 5, 7] = 0.0                                       # an exact zero (mfg_ac2.py:244)
    ref = O.rollout_frozen(pi0.astype(np.float64), 8.64, 0.05, 1e4, y.astype(np.float64), w=w)
    out = eng.rollout(T_(pi0, dev, torch.float32), 8.64, 0.05, 1e4, T, w=T_(w, dev, torch.float64),
                      noise_y=T_(y, dev, torch.float32), outputs=ALL_OUT, want_acc=True, variant="generic")
    for k in ("states", "alpha"):
        np.testing.assert_allclose(N_(out[k]), ref[k], rtol=1e-5, err_msg=k)
    np.testing.assert_allclose(N_(out["alpha_deriv"]), ref["alpha_deriv"], rtol=1e-5, atol=3e-8)
    np.testing.assert_allclose(N_(out["actions"]), ref["actions"], rtol=1e-5, atol=1e-30)
    np.testing.assert_allclose(N_(out["pi_final"]), ref["states"][-1], rtol=1e-5)
    # g sums d^2 terms of mixed sign: bound relative to the sum of their magnitudes (~ d for these inputs)
    np.testing.assert_allclose(N_(out["grads"]), ref["grads"], rtol=1e-5, atol=2e-6 * d)
    v = np.abs(O.features(ref["states"]) @ w)
    scale = np.abs(ref["rewards"]) + v[1:] + v[:-1]
    assert np.all(np.abs(N_(out["deltas"]) - ref["deltas"]) <= 1e-6 * scale)
    assert np.all(np.abs(N_(out["rewards"]) - ref["rewards"]) <= 1e-6 * np.maximum(scale, 1e-3))
    acc = N_(out["acc"])
    wscale = np.sum(np.abs(ref["deltas"])[..., None] * np.abs(O.features(ref["states"][:-1])), axis=(0, 1))
    assert np.all(np.abs(acc[1:1 + F] - ref["G_w"]) <= 1e-5 * wscale + 1e-12)
    np.testing.assert_allclose(acc[1 + F], ref["R"], rtol=1e-5, atol=1e-9)


def test_random_regimes_invariants_and_variant_agreement(dev):
    """30 random parameter regimes (theta 1..30, shift 0..0.6, alpha_scale 10..1e5: shapes from 1e-4 -- boost and redo
    paths -- to 1e6, d in {15, 16, 20, 21, 32, 64}, ragged B, odd / even T): every output finite, rows of P on the simplex,
    mass conserved, and -- the draws being keyed by (seed, population, step, row, pair) -- the v2 kernel and the wide
    kernel produce the same trajectories wherever both exist."""
    rng = np.random.RandomState(2024)
    for trial in range(30):
        d = int(rng.choice([15, 16, 20, 21, 32, 64]))
        B, T = int(rng.randint(1, 70)), int(rng.randint(1, 9))
        theta, shift = float(rng.uniform(1.0, 30.0)), float(rng.uniform(0.0, 0.6))
        scale = float(rng.choice([10.0, 100.0, 1e4, 1e5]))
        conc = float(rng.choice([0.05, 0.3, 1.0]))                       # sparse .. flat start states
        pi0 = T_(rng.dirichlet(np.ones(d) * conc, size=B), dev, torch.float32)
        w = T_(rng.rand(O.num_features(d)), dev, torch.float64)
        reward = ("ac2", "synthetic")[trial % 2]                        # mfg_ac2.py:257-287 / mfg_synthetic.py:249-265
        kw = dict(w=w, seed=trial, reward=reward, outputs=("states", "actions", "rewards", "grads", "deltas"),
                  want_acc=True)
        a = eng.rollout(pi0, theta, shift, scale, T, **kw)
        tag = "trial %d: d=%d B=%d T=%d theta=%.2f shift=%.2f scale=%g" % (trial, d, B, T, theta, shift, scale)
        for k, v in a.items():
            assert torch.isfinite(v).all(), tag + " " + k
        P, S = a["actions"].double(), a["states"].double()
        assert torch.all(P > 0), tag
        assert float((P.sum(-1) - 1).abs().max()) <= 2e-6, tag
        assert float((S.sum(-1) - S[0].sum(-1)).abs().max()) <= 2e-6, tag
        assert float((torch.einsum("tbi,tbij->tbj", S[:-1], P) - S[1:]).abs().max()) <= 3e-7, tag
        if d in (15, 16, 20, 21, 32):
            # the TRAIN specialisation (no per-step stream: one merged reduction per step) sums the same things
            tr = eng.rollout(pi0, theta, shift, scale, T, w=w, seed=trial, reward=reward, outputs=(), want_acc=True)
            sc = float(a["deltas"].double().abs().sum()) * max(1.0, float(a["grads"].double().abs().max()))
            assert float((tr["acc"] - a["acc"]).abs().max()) <= 1e-9 * max(sc, 1e-30) + 1e-12, tag
            g = eng.rollout(pi0, theta, shift, scale, T, variant="generic", **kw)
            np.testing.assert_allclose(N_(a["states"]), N_(g["states"]), rtol=3e-5, atol=1e-8, err_msg=tag)
            np.testing.assert_allclose(N_(a["rewards"]), N_(g["rewards"]), rtol=1e-4, atol=1e-7, err_msg=tag)


def test_random_regimes_learners_float_vs_double(dev):
    """The v2 learners (float streams) against the first-generation learners in float64 on the SAME Philox draws, in
    12 random regimes (theta, shift, alpha_scale, step sizes, both reward / discount flavours): the per-step online
    updates of (theta, w) agree to float accuracy after several episodes, and stay finite."""
    rng = np.random.RandomState(77)
    for trial in range(12):
        d = int(rng.choice([15, 16]))
        L, E, T = int(rng.randint(1, 40)), int(rng.randint(1, 4)), int(rng.randint(1, 16))
        theta0 = rng.uniform(2.0, 20.0, size=L)
        shifts = rng.uniform(0.0, 0.5, size=L)
        scale = float(rng.choice([100.0, 1e4, 1e5]))
        mat = rng.dirichlet(np.ones(d) * float(rng.choice([0.2, 1.0])), size=11)
        w0 = rng.rand(L, O.num_features(d))
        reward, discount = [("ac2", "step"), ("synthetic", "cumulative")][trial % 2]
        res = {}
        for dt in (torch.float32, torch.float64):
            th = torch.as_tensor(theta0, dtype=torch.float64, device=dev).clone()
            w = torch.as_tensor(w0, dtype=torch.float64, device=dev).clone()
            eng.learners(th, w, T_(np.float32(mat), dev, dt), E, T,
                         shift=torch.as_tensor(shifts, dtype=torch.float64, device=dev), alpha_scale=scale, episode0=1,
                         gamma=0.97, lr_critic=0.1, lr_actor=0.01, reward=reward, discount=discount, seed=1000 + trial)
            res[dt] = (N_(th), N_(w))
        tag = "trial %d: d=%d L=%d E=%d T=%d scale=%g %s/%s" % (trial, d, L, E, T, scale, reward, discount)
        assert np.isfinite(res[torch.float32][0]).all() and np.isfinite(res[torch.float32][1]).all(), tag
        np.testing.assert_allclose(res[torch.float32][0], res[torch.float64][0], rtol=5e-6, err_msg=tag)
        np.testing.assert_allclose(res[torch.float32][1], res[torch.float64][1], rtol=5e-5, atol=5e-6, err_msg=tag)
