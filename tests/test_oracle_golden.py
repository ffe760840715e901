"""The oracle against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from /root/reference) -- this is what pins the oracle."""
import numpy as np

from oracle import mfg_oracle as O


def test_kat_reward_value_features(kat):
    # test2.py:46-56, 73-88 inputs; SURVEY App. B outputs
    assert np.isclose(O.reward_ac2(kat["P3"], kat["pi3"]), kat["reward3"], rtol=1e-14)
    assert np.isclose(kat["reward3"], -39.07, rtol=1e-12)
    assert np.isclose(O.cost_ac1(kat["P3"], kat["pi3"]), kat["cost3_ac1"], rtol=1e-14)
    assert np.isclose(O.value(kat["pi3"], np.ones(10)), kat["value3_w1"], rtol=1e-14)
    assert np.isclose(kat["value3_w1"], 2.77, rtol=1e-12)
    np.testing.assert_array_equal(O.features(np.array([2.0, 3.0, 5.0])), kat["features_235"])
    np.testing.assert_array_equal(kat["features_235"], [4, 6, 10, 9, 15, 25, 2, 3, 5, 1])
    assert np.isclose(O.reward_synthetic(kat["P3_rowstochastic"], kat["pi3"]),
                      kat["reward3_synthetic"], rtol=1e-14)


def test_kat_d4_policy(kat):
    # test2.py:6,16,105-121
    th, sh, sc = float(kat["d4_theta"]), float(kat["d4_shift"]), float(kat["d4_alpha_scale"])
    P, alpha, deriv, _ = O.sample_action(kat["d4_pi"], th, sh, sc, y=kat["d4_y"])
    np.testing.assert_allclose(alpha, kat["d4_alpha"], rtol=1e-14)
    np.testing.assert_allclose(deriv, kat["d4_alpha_deriv"], rtol=1e-14)
    np.testing.assert_allclose(P, kat["d4_P"], rtol=1e-15)
    np.testing.assert_allclose(O.mean_field_step(P, kat["d4_pi"]), kat["d4_pi_next"], rtol=1e-14)
    assert np.isclose(O.reward_ac2(P, kat["d4_pi"]), kat["d4_reward"], rtol=1e-13)
    g = O.log_policy_gradient(alpha, deriv, P)
    for k in ("d4_grad_basic", "d4_grad_loop", "d4_grad_vectorized"):   # three formulations agree
        assert np.isclose(g, kat[k], rtol=1e-13)
    np.testing.assert_allclose(O.features(kat["d4_pi"]), kat["d4_features"], rtol=1e-15)
    assert np.isclose(O.reward_synthetic(P, kat["d4_pi"]), kat["d4_reward_synthetic"], rtol=1e-14)


def test_kat_d4_global_rng_stream(kat):
    """Same seed + same call pattern reproduces the reference's Gamma stream."""
    np.random.seed(0)
    P, *_ , y = O.sample_action(kat["d4_pi"], float(kat["d4_theta"]), float(kat["d4_shift"]),
                                float(kat["d4_alpha_scale"]))
    np.testing.assert_array_equal(y, kat["d4_y"])
    np.testing.assert_allclose(P, kat["d4_P"], rtol=1e-15)


def test_train_trace_serial_learner(trace):
    """Config 1: the reference's train() replayed with its own draws."""
    E, T = int(trace["episodes"]), int(trace["T"])
    noise = O.InjectedNoise(trace["start_rows"], trace["y"])
    theta, w, info = O.train_serial(
        trace["mat_pi0"], float(trace["theta0"]), trace["w0"], float(trace["shift"]),
        float(trace["alpha_scale"]), E, gamma=float(trace["gamma"]), constant=False,
        lr_critic=float(trace["lr_critic"]), lr_actor=float(trace["lr_actor"]),
        flavour="mfg_ac2", reward="ac2", noise=noise, num_steps=T, trace=True)
    steps = info["steps"]
    for e in range(E):
        for t in range(T):
            s = steps[e * T + t]
            np.testing.assert_allclose(s["pi"], trace["pi"][e, t], rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(s["P"], trace["P"][e, t], rtol=1e-12)
            np.testing.assert_allclose(s["pi_next"], trace["pi_next"][e, t], rtol=1e-12)
            assert np.isclose(s["reward"], trace["reward"][e, t], rtol=1e-9, atol=1e-18)
            assert np.isclose(s["grad"], trace["grad"][e, t], rtol=1e-10)
            assert np.isclose(s["delta"], trace["delta"][e, t], rtol=1e-9, atol=1e-15)
            assert np.isclose(s["theta"], trace["theta_after"][e, t], rtol=1e-12)
            np.testing.assert_allclose(s["w"], trace["w_after"][e, t], rtol=1e-12)
    assert np.isclose(theta, float(trace["theta_final"]), rtol=1e-12)
    np.testing.assert_allclose(w, trace["w_final"], rtol=1e-12)


def test_train_trace_global_rng(trace):
    """Seeding NumPy like make_golden did reproduces the whole reference run
    (init_w draw, start rows, Gamma stream) through the oracle alone."""
    d = int(trace["d"])
    np.random.seed(0)
    w0 = np.random.rand(O.num_features(d), 1).reshape(-1)          # mfg_ac2.py:176
    np.testing.assert_array_equal(w0, trace["w0"])
    theta, w, _ = O.train_serial(trace["mat_pi0"], float(trace["theta0"]), w0, float(trace["shift"]),
                                 float(trace["alpha_scale"]), int(trace["episodes"]),
                                 lr_critic=0.1, lr_actor=0.1, flavour="mfg_ac2")
    assert np.isclose(theta, float(trace["theta_final"]), rtol=1e-12)
    np.testing.assert_allclose(w, trace["w_final"], rtol=1e-12)


def test_frozen_rollout_matches_trace_first_steps(trace):
    """rollout_frozen with (theta0, w0) frozen equals the reference at step 0 of
    episode 0 (the only step where nothing has been updated yet) and keeps the
    recurrence state_{t+1} = action_t^T state_t (test_acirl.py:43-47)."""
    y = trace["y"][0][:, None]                      # [T,1,d,d]
    pi0 = trace["mat_pi0"][int(trace["start_rows"][0])][None]
    out = O.rollout_frozen(pi0, float(trace["theta0"]), float(trace["shift"]),
                           float(trace["alpha_scale"]), y, w=trace["w0"])
    np.testing.assert_allclose(out["actions"][0, 0], trace["P"][0, 0], rtol=1e-12)
    assert np.isclose(out["deltas"][0, 0], trace["delta"][0, 0], rtol=1e-9)
    assert np.isclose(out["grads"][0, 0], trace["grad"][0, 0], rtol=1e-10)
    for t in range(y.shape[0]):
        np.testing.assert_allclose(out["states"][t + 1, 0],
                                   out["actions"][t, 0].T @ out["states"][t, 0], rtol=1e-13)
        np.testing.assert_allclose(out["actions"][t, 0].sum(-1), 1.0, rtol=1e-13)   # test2.py:26


def test_generate_trajectory(traj15):
    noise = O.InjectedNoise(np.zeros(1, int), traj15["y"][None])
    noise.start_index(1)
    out = O.generate_trajectory(traj15["pi0"], 16, float(traj15["theta"]), float(traj15["shift"]),
                                float(traj15["alpha_scale"]), noise)
    np.testing.assert_allclose(out, traj15["trajectory"], rtol=1e-12)
    # start rows are not renormalised (quirk C.4): the mass is preserved, not 1
    np.testing.assert_allclose(out.sum(1), traj15["pi0"].sum(), rtol=1e-12)


def test_forward_d47(fwd47):
    pi = fwd47["states"][0]
    for t in range(3):
        o = O.transition(pi, float(fwd47["theta"]), float(fwd47["shift"]), float(fwd47["alpha_scale"]),
                         fwd47["y"][t])
        np.testing.assert_allclose(o["P"], fwd47["actions"][t], rtol=1e-12)
        assert np.isclose(o["reward"], fwd47["reward"][t], rtol=1e-9)
        assert np.isclose(o["grad"], fwd47["grad"][t], rtol=1e-10)
        pi = o["pi_next"]
        np.testing.assert_allclose(pi, fwd47["states"][t + 1], rtol=1e-12)


def test_float32_cancellation_is_where_the_survey_says(trace):
    """SURVEY section 7 hard part 2: plain float32 is fine for pi', V, g but not for r, delta."""
    y = trace["y"].reshape(-1, 1, 15, 15)[:15]
    pi0 = trace["mat_pi0"][:1]
    a64 = O.rollout_frozen(pi0, 8.86349, 0.16, 12000.0, y, w=trace["w0"])
    a32 = O.rollout_frozen(pi0.astype(np.float32), np.float32(8.86349), np.float32(0.16),
                           np.float32(12000.0), y.astype(np.float32), w=trace["w0"], dtype=np.float32)
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    assert rel(a32["states"], a64["states"]) < 5e-6
    assert rel(a32["grads"], a64["grads"]) < 2e-5
    # deltas lose several digits in float32 -- the CUDA path accumulates these in float64
    assert rel(a32["deltas"], a64["deltas"]) > 1e-6


def test_metrics_oracle_against_reference_evaluate(evalm):
    """jsd / trajectory_metrics against the reference's own JSD and evaluate() (mfg_ac2.py:546-563, 595-670)."""
    assert np.isclose(O.jsd(evalm["jsd_P"], evalm["jsd_Q"]), float(evalm["jsd_value"]), rtol=1e-13)
    emp, y = evalm["empirical"], evalm["y"]
    gen = []
    for k in range(emp.shape[0]):
        noise = O.InjectedNoise(start_rows=np.zeros(1, int), y=y[k][None])
        noise.start_index(1)
        gen.append(O.generate_trajectory(emp[k, 0], 16, float(evalm["theta"]), float(evalm["shift"]),
                                         float(evalm["alpha_scale"]), noise))
    res = [np.mean(a) for a in O.trajectory_metrics(np.stack(gen), emp)]
    np.testing.assert_allclose(res, evalm["result"], rtol=1e-12)


def test_synthetic_check_oracle_against_reference():
    """mfg_synthetic.evaluate_synthetic / evaluate_synthetic_JSD (mfg_synthetic.py:741-899) on the actions the
    reference itself sampled: the oracle reproduces the (mean, std) both calls returned."""
    from conftest import load_golden
    g = load_golden("synthetic_check.npz")
    l1 = np.concatenate([O.synthetic_check(a)[0] for a in g["actions_l1"]])
    js = np.concatenate([O.synthetic_check(a)[1] for a in g["actions_jsd"]])
    np.testing.assert_allclose([l1.mean(), l1.std()], g["l1_mean_std"], rtol=1e-12)
    np.testing.assert_allclose([js.mean(), js.std()], g["jsd_mean_std"], rtol=1e-12)
