"""The drop-in ``mfg_ac2.actor_critic`` class: the reference's hand tests (test2.py) turned into
assertions, run through the public NumPy-in / NumPy-out API on the GPU."""
import numpy as np
import pytest
import torch

from oracle import mfg_oracle as O

pytestmark = pytest.mark.gpu

mfg_ac2 = pytest.importorskip("discrete_mean_field_game_b200.mfg_ac2")


@pytest.fixture(scope="module")
def start_states():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return O.synthetic_start_states(n_rows=21, n_cols=20, d=20, seed=0)


def make(start_states, **kw):
    kw.setdefault("dtype", "float64")
    kw.setdefault("seed", 1)
    return mfg_ac2.actor_critic(mat_pi0=start_states, **kw)


def test_constructor_surface(start_states):
    ac = make(start_states)                                   # mfg_ac2.py:25 defaults
    assert (ac.theta, ac.shift, ac.alpha_scale, ac.d) == (8.86349, 0.16, 12000, 21)
    assert ac.w.shape == (21 * 22 // 2 + 21 + 1, 1) and ac.w.min() >= 0 and ac.w.max() < 1
    assert ac.mat_pi0.shape == (21, 20) and ac.num_start_samples == 21
    assert ac.mat_alpha.shape == (21, 21) and ac.mat_alpha_deriv.shape == (21, 21)


def test_init_pi0_reads_reference_format(tmp_path, start_states):
    d = tmp_path / "train_normalized_round2"
    d.mkdir()
    for k, row in enumerate(start_states, start=1):
        lines = [" ".join("%.3e" % v for v in row)] + [" ".join("%.3e" % v for v in row[::-1])] * 15
        (d / ("trend_distribution_day%d.csv" % k)).write_text("\n".join(lines) + "\n")
    ac = mfg_ac2.actor_critic(d=15, path_to_dir=str(d), dtype="float64", seed=0)
    np.testing.assert_array_equal(ac.mat_pi0, start_states[:, :15])


def test_action(start_states, kat):
    """test2.py:14-32: rows of P and the new state keep their mass; with the reference's draws
    injected the reference's P comes back."""
    ac = make(start_states, theta=10, shift=0.4, d=4)
    pi = np.array([0.7, 0.09, 0.01, 0.2])
    P = ac.sample_action(pi)
    assert P.shape == (4, 4) and np.all(P > 0)
    np.testing.assert_allclose(P.sum(1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(P.T.dot(pi).sum(), 1.0, rtol=1e-12)
    P2 = ac.sample_action(pi)
    assert not np.allclose(P, P2)                              # the stream advances
    Pg = ac.sample_action(pi, y=kat["d4_y"])
    np.testing.assert_allclose(Pg, kat["d4_P"], rtol=1e-12)
    np.testing.assert_allclose(ac.mat_alpha, kat["d4_alpha"], rtol=1e-10)
    np.testing.assert_allclose(ac.mat_alpha_deriv, kat["d4_alpha_deriv"], rtol=1e-10)
    np.testing.assert_allclose(ac.step(Pg, pi), kat["d4_pi_next"], rtol=1e-12)


def test_reward(start_states, kat):
    """test2.py:46-56: the fixed (non-stochastic) 3x3 input -> -39.07."""
    ac = make(start_states)
    P = np.array([[1, 3, 3], [4, 5, 6], [7, 8, 9]])
    r = ac.calc_reward(P, np.array([0.1, 0.2, 0.7]), 3)
    assert r.shape == (1,)
    np.testing.assert_allclose(r[0], -39.07, rtol=1e-12)
    np.testing.assert_allclose(r[0], kat["reward3"], rtol=1e-12)
    ac32 = make(start_states, dtype="float32")
    np.testing.assert_allclose(ac32.calc_reward(P, np.array([0.1, 0.2, 0.7]), 3)[0], -39.07, rtol=1e-6)


def test_value_and_features(start_states, kat):
    """test2.py:73-88: d=3, w=1 -> 2.77; feature order of the code, not of the docstring."""
    ac = make(start_states, d=3)
    ac.w = np.ones(10)
    np.testing.assert_allclose(ac.calc_value(np.array([0.1, 0.2, 0.7])), 2.77, rtol=1e-13)
    np.testing.assert_array_equal(ac.calc_features(np.array([2.0, 3.0, 5.0])), kat["features_235"])
    ac4 = make(start_states, d=4)
    np.testing.assert_allclose(ac4.calc_features(kat["d4_pi"]), kat["d4_features"], rtol=1e-15)


def test_gradient_first(start_states, kat):
    """test2.py:105-121: the three gradient entry points agree (and equal the reference's value)."""
    ac = make(start_states, theta=10, shift=0.4, d=4)
    pi = kat["d4_pi"]
    P = ac.sample_action(pi, y=kat["d4_y"])
    P_before = P.copy()
    g1, g2, g3 = ac.calc_gradient_basic(P, pi), ac.calc_gradient(P, pi), ac.calc_gradient_vectorized(P, pi)
    assert g1 == g2 == g3
    np.testing.assert_allclose(g3, kat["d4_grad_vectorized"], rtol=1e-10)
    np.testing.assert_array_equal(P, P_before)                 # inputs are not mutated (quirk A.11)
    np.testing.assert_allclose(ac.calc_reward(P, pi, 4)[0], kat["d4_reward"], rtol=1e-10)


def test_forward_d47(start_states, fwd47):
    """test2.py:203-224: multi-step rollout at d = 47 (generic kernel)."""
    ac = make(start_states, d=4)
    ac.d = 47
    pi = fwd47["states"][0]
    for t in range(3):
        P = ac.sample_action(pi, y=fwd47["y"][t])
        np.testing.assert_allclose(P, fwd47["actions"][t], rtol=1e-11)
        pi = np.transpose(P).dot(pi)
    np.testing.assert_allclose(pi, fwd47["states"][3], rtol=1e-11)
    traj = ac.generate_trajectory(fwd47["states"][0], 4, y=fwd47["y"])
    np.testing.assert_allclose(traj, fwd47["states"], rtol=1e-11)
    free = ac.generate_trajectory(fwd47["states"][0], 16)
    assert free.shape == (16, 47)
    np.testing.assert_allclose(free.sum(1), 1.0, rtol=1e-10)


def test_generate_trajectory_d15(start_states, traj15):
    ac = make(start_states, d=15)
    out = ac.generate_trajectory(traj15["pi0"], 16, y=traj15["y"])
    np.testing.assert_allclose(out, traj15["trajectory"], rtol=1e-11)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_train_replays_reference(start_states, trace, dtype, capsys):
    """Config 1 through the class: mfg_ac2.py:831-836 with the reference's recorded draws."""
    ac = mfg_ac2.actor_critic(theta=float(trace["theta0"]), shift=0.16, alpha_scale=12000, d=15,
                              mat_pi0=trace["mat_pi0"], dtype=dtype, seed=0)
    ac.w = trace["w0"].reshape(-1, 1).copy()
    ac.train(num_episodes=3, gamma=1, constant=0, lr_critic=0.1, lr_actor=0.1,
             start_rows=trace["start_rows"], noise_y=trace["y"])
    rt = 1e-11 if dtype == "float64" else 1e-6
    np.testing.assert_allclose(ac.theta, float(trace["theta_final"]), rtol=rt)
    np.testing.assert_allclose(ac.w.ravel(), trace["w_final"], rtol=10 * rt, atol=1e-7 if dtype == "float32" else 0)
    assert ac.w.shape == (136, 1)
    assert "Average reward during previous 100 episodes" in capsys.readouterr().out


def test_train_generic_d_matches_oracle(start_states):
    """d = 21 (the reference's default) in float64 has no fused learner kernel: host-driven per-step path
    (float32 streams run learners_v2_kernel<21, 32>, tested in test_rollout_gpu.py)."""
    rng = np.random.RandomState(2)
    d, E, T = 21, 2, 15
    mat = O.synthetic_start_states(n_rows=21, n_cols=30, d=21, seed=4)
    ac = mfg_ac2.actor_critic(d=d, mat_pi0=mat, dtype="float64", seed=3)
    w0 = ac.w.ravel().copy()
    start = rng.randint(0, 21, size=E)
    y = rng.gamma(shape=200.0, size=(E, T, d, d))
    ac.train(num_episodes=E, lr_critic=0.1, lr_actor=0.001, start_rows=start, noise_y=y, verbose=False)
    th, w, _ = O.train_serial(mat, 8.86349, w0, 0.16, 12000, E, lr_critic=0.1, lr_actor=0.001,
                              flavour="mfg_ac2", noise=O.InjectedNoise(start, y))
    np.testing.assert_allclose(ac.theta, th, rtol=1e-11)
    np.testing.assert_allclose(ac.w.ravel(), w, rtol=1e-10)


def test_train_philox_is_reproducible_and_learns_something(start_states):
    a = mfg_ac2.actor_critic(d=15, mat_pi0=start_states, dtype="float32", seed=42)
    b = mfg_ac2.actor_critic(d=15, mat_pi0=start_states, dtype="float32", seed=42)
    b.w = a.w.copy()
    a.train(num_episodes=30, lr_critic=0.1, lr_actor=0.1, consecutive=10, verbose=False)
    b.train(num_episodes=30, lr_critic=0.1, lr_actor=0.1, consecutive=7, verbose=False)   # chunking is invisible
    assert a.theta == b.theta and np.array_equal(a.w, b.w)
    assert a.theta != 8.86349 and np.isfinite(a.theta)
    th1 = a.theta
    a.train(num_episodes=5, verbose=False)                      # a second call continues the noise stream
    assert a.theta != th1


def test_train_batch_per_step_reduces_to_train_at_B1(start_states, trace):
    """SURVEY 7 hard part 1: the per_step batched update IS the reference's update at B = 1."""
    ac = mfg_ac2.actor_critic(d=15, mat_pi0=trace["mat_pi0"], dtype="float64", seed=9)
    w0 = ac.w.copy()
    pi0 = trace["mat_pi0"][4:5]
    ac.train_batch(pi0, num_episodes=2, T=15, lr_critic=0.1, lr_actor=0.1, update="per_step")
    th_b, w_b = ac.theta, ac.w.ravel().copy()
    # replay through the serial learner with the same Philox draws: record them via rollout_batch
    ref = mfg_ac2.actor_critic(d=15, mat_pi0=trace["mat_pi0"], dtype="float64", seed=9)
    ref.w = w0.copy()
    theta, w = 8.86349, w0.ravel().copy()
    import math
    from discrete_mean_field_game_b200 import engine
    for e in range(2):
        pi = pi0[0]
        for t in range(15):
            ref.theta = theta
            out = engine.rollout(ref._dev(pi[None]), theta, 0.16, 12000, 1, seed=9, step_offset=e * 15 + t,
                                 reward="none", outputs=("actions",))
            P = out["actions"][0, 0].cpu().numpy()
            alpha, deriv = O.policy_alpha(pi, theta, 0.16)
            pn = O.mean_field_step(P, pi)
            delta = O.reward_ac2(P, pi) + O.features(pn) @ w - O.features(pi) @ w
            g = O.log_policy_gradient(alpha, deriv, P)
            w = w + O.critic_lr(e, 0.1, False) * delta * O.features(pi)
            theta = theta + O.actor_lr(e, 0.1, False) * delta * g
            pi = pn
    np.testing.assert_allclose(th_b, theta, rtol=1e-10)
    np.testing.assert_allclose(w_b, w, rtol=1e-10)


def test_train_batch_per_episode_matches_oracle(start_states):
    """per_episode: frozen parameters, one batch-mean update per episode."""
    rng = np.random.RandomState(0)
    B, T, d = 64, 15, 15
    pi0 = rng.dirichlet(np.ones(d), size=B)
    ac = mfg_ac2.actor_critic(d=d, mat_pi0=start_states, dtype="float64", seed=5)
    w0 = ac.w.ravel().copy()
    rec = ac.rollout_batch(pi0, T=T, record=True, seed=5)           # same draws as episode 0 of train_batch
    res = ac.train_batch(pi0, num_episodes=1, T=T, lr_critic=0.1, lr_actor=0.01, update="per_episode", seed=5)
    alpha, deriv = O.policy_alpha(rec["states"][:-1], 8.86349, 0.16)
    g = O.log_policy_gradient(alpha, deriv, rec["actions"])
    r = O.reward_ac2(rec["actions"], rec["states"][:-1])
    phi = O.features(rec["states"])
    v = phi @ w0
    delta = r + v[1:] - v[:-1]
    np.testing.assert_allclose(rec["rewards"], r, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(rec["deltas"], delta, rtol=1e-8, atol=1e-14)
    np.testing.assert_allclose(rec["grads"], g, rtol=1e-9)
    th = 8.86349 + O.actor_lr(0, 0.01, False) / B * np.sum(delta * g)
    w = w0 + O.critic_lr(0, 0.1, False) / B * np.einsum("tb,tbf->f", delta, phi[:-1])
    np.testing.assert_allclose(res["theta"], th, rtol=1e-11)
    np.testing.assert_allclose(ac.w.ravel(), w, rtol=1e-10)
    np.testing.assert_allclose(res["mean_reward"][0], r.sum() / B, rtol=1e-9)


def test_train_batch_default_d21_float32_matches_oracle(start_states):
    """The reference's default constructor (d = 21, mfg_ac2.py:25) on float streams: the batched train step runs the
    32-lane v2 kernel (TD error and critic Gram in the kernel).  Its update equals the oracle's batch-mean update
    evaluated in float64 on the recorded episode (same Philox draws); tolerances are relative to the sum of the
    magnitudes of the summed terms (float32 streams, 441 terms of mixed sign per gradient)."""
    rng = np.random.RandomState(3)
    B, T = 203, 7                                                  # ragged tile, T ends inside a Gram flush group
    ac = mfg_ac2.actor_critic(mat_pi0=start_states, dtype="float32", seed=5)
    d = ac.d
    assert d == 21
    pi0 = rng.dirichlet(np.ones(d), size=B)
    w0 = ac.w.ravel().astype(np.float64).copy()
    rec = ac.rollout_batch(pi0, T=T, record=True, seed=5)
    res = ac.train_batch(pi0, num_episodes=1, T=T, lr_critic=0.1, lr_actor=0.01, update="per_episode", seed=5)
    S, P = np.float64(rec["states"]), np.float64(rec["actions"])
    alpha, deriv = O.policy_alpha(S[:-1], 8.86349, 0.16)
    g = O.log_policy_gradient(alpha, deriv, P)
    r = O.reward_ac2(P, S[:-1])
    phi = O.features(S)
    v = phi @ w0
    delta = r + v[1:] - v[:-1]
    la, lc = O.actor_lr(0, 0.01, False) / B, O.critic_lr(0, 0.1, False) / B
    th = 8.86349 + la * np.sum(delta * g)
    w = w0 + lc * np.einsum("tb,tbf->f", delta, phi[:-1])
    assert abs(res["theta"] - th) <= 1e-4 * la * np.sum(np.abs(delta * g)) + 1e-12, (res["theta"], th)
    wscale = lc * np.einsum("tb,tbf->f", np.abs(delta), np.abs(phi[:-1]))
    err = np.abs(ac.w.ravel().astype(np.float64) - w)
    assert np.all(err <= 1e-5 * wscale + 1e-12), np.max(err / (wscale + 1e-300))
    np.testing.assert_allclose(res["mean_reward"][0], r.sum() / B, rtol=1e-5)


def test_mfg_synthetic_dropin(start_states):
    """mfg_synthetic.actor_critic: synthetic reward (mfg_synthetic.py:249-265) and the (shift, theta0) sweep of
    its __main__ (:902-925) as independent learners -- each learner equals a separate train() run."""
    from discrete_mean_field_game_b200 import mfg_synthetic
    np.random.seed(0)
    ac = mfg_synthetic.actor_critic(theta=2.6, shift=0.02, alpha_scale=10000, d=15, mat_pi0=start_states,
                                    dtype="float64", seed=3)
    rng = np.random.RandomState(1)
    P = rng.dirichlet(np.ones(15), size=15)
    pi = rng.dirichlet(np.ones(15))
    np.testing.assert_allclose(ac.calc_reward(P, pi, 15)[0], -0.5 * pi.dot((P * P).sum(1)), rtol=1e-12)
    np.random.seed(5)
    res = ac.sweep(shifts=[0.0, 0.02], thetas=[1.0, 2.5, 4.0], num_episodes=6)
    assert res.shape == (6, 3) and np.all(np.isfinite(res))
    np.testing.assert_array_equal(res[:, 0], [0, 0, 0, 0.02, 0.02, 0.02])
    assert np.all(res[:, 2] != res[:, 1])                    # every learner moved
    # learner 4 (shift 0.02, theta0 2.5) == an independent train() with the same w0 / seed / learner id
    np.random.seed(5)
    w0 = np.random.rand(6, 136)
    import torch
    from discrete_mean_field_game_b200 import engine
    th = torch.tensor([2.5], dtype=torch.float64, device=ac.device)
    w = torch.as_tensor(w0[4:5].copy(), device=ac.device)
    engine.learners(th, w, ac._dev(ac.mat_pi0), 6, 15, shift=0.02, alpha_scale=10000.0, lr_critic=0.1,
                    lr_actor=0.001, constant=True, reward="synthetic", seed=3, learner_offset=4,
                    want_total_reward=False)
    np.testing.assert_allclose(res[4, 2], float(th[0]), rtol=1e-12)


def test_evaluate_and_jsd_match_reference(start_states, evalm, tmp_path):
    """actor_critic.evaluate / JSD / gridsearch (mfg_ac2.py:546-689) with the reference's own draws."""
    d = int(evalm["d"])
    ac = mfg_ac2.actor_critic(d=d, mat_pi0=start_states, dtype="float64", seed=1)
    np.testing.assert_allclose(ac.JSD(evalm["jsd_P"], evalm["jsd_Q"]), float(evalm["jsd_value"]), rtol=1e-12)
    out = tmp_path / "eval" / "res.csv"
    res = ac.evaluate(theta=float(evalm["theta"]), shift=float(evalm["shift"]), alpha_scale=float(evalm["alpha_scale"]),
                      d=d, episode_length=16, outfile=str(out), write_header=1, empirical=evalm["empirical"],
                      y=evalm["y"])
    np.testing.assert_allclose(res, evalm["result"], rtol=1e-10)
    assert out.read_text() == str(evalm["csv"])                       # same header and formatted line
    # reading the test days from files (sorted order), float32 streams, Philox noise
    indir = tmp_path / "days"
    indir.mkdir()
    for k, m in enumerate(evalm["empirical"]):
        np.savetxt(indir / ("trend_distribution_day%d.csv" % (22 + k)), m, fmt="%.6e", delimiter=" ")
    import os
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ac32 = mfg_ac2.actor_critic(d=d, mat_pi0=start_states, dtype="float32", seed=2)
        r1 = ac32.evaluate(theta=8.0, shift=0.16, alpha_scale=12000, d=d, indir="days", outfile=None)
        assert all(np.isfinite(r1)) and abs(r1[0] - float(evalm["result"][0])) < 0.1
        best = ac32.gridsearch([7.0, 8.0], [0.1, 0.16], [12000], "days", str(tmp_path / "grid.csv"), verbose=False)
        assert len(best) == 4 and all(b[1] in (7.0, 8.0) and b[2] in (0.1, 0.16) for b in best)
        assert len(open(tmp_path / "grid.csv").read().strip().splitlines()) == 4
    finally:
        os.chdir(cwd)


def test_synthetic_analytic_check_matches_reference():
    """evaluate_synthetic / evaluate_synthetic_JSD (mfg_synthetic.py:741-899): the GPU backward-equation check on
    the reference's own sampled actions returns the reference's (mean, std); on its own rollout it matches the
    oracle applied to the recorded actions; generate_trajectory returns (states, actions) like the reference."""
    from conftest import load_golden
    from discrete_mean_field_game_b200 import engine, mfg_synthetic
    g = load_golden("synthetic_check.npz")
    d = int(g["d"])
    ac = mfg_synthetic.actor_critic(theta=float(g["theta"]), shift=float(g["shift"]), alpha_scale=float(g["alpha_scale"]),
                                    d=d, mat_pi0=g["mat_pi0"], seed=3)
    np.testing.assert_allclose(ac.evaluate_synthetic(1, 3, actions=g["actions_l1"]), g["l1_mean_std"], rtol=1e-10)
    np.testing.assert_allclose(ac.evaluate_synthetic_JSD(1, 3, actions=g["actions_jsd"]), g["jsd_mean_std"], rtol=1e-10)
    np.testing.assert_allclose(ac.calc_reward_vector(g["actions_l1"][0, 0]),
                               -0.5 * (g["actions_l1"][0, 0] ** 2).sum(1), rtol=1e-14)
    # own rollout (float streams): kernel vs oracle on the recorded actions
    pi0 = torch.as_tensor(g["mat_pi0"], dtype=torch.float32, device=ac.device)
    acts = engine.rollout(pi0, ac.theta, ac.shift, ac.alpha_scale, 15, reward="none", seed=5, outputs=("actions",))["actions"]
    l1, js = engine.synthetic_check(acts)
    A = acts.double().cpu().numpy()
    for b in range(A.shape[1]):
        rl1, rjs = O.synthetic_check(A[:, b])
        np.testing.assert_allclose(l1[b].cpu().numpy(), rl1, rtol=1e-10)
        np.testing.assert_allclose(js[b].cpu().numpy(), rjs, rtol=1e-9)
    m, s_ = ac.evaluate_synthetic(1, 3)
    assert np.isfinite(m) and np.isfinite(s_)
    traj, actions = ac.generate_trajectory(g["mat_pi0"][0], 16)
    assert traj.shape == (16, d) and actions.shape == (15, d, d)
    np.testing.assert_allclose(np.einsum("ti,tij->tj", traj[:-1], actions), traj[1:], atol=3e-7)



def test_train_batch_pipelined_inputs_and_history(start_states):
    """Per-episode host inputs (list / callable) through the double-buffered copy pipeline give the same
    trajectory of (theta, w) as one call per episode, and history=True returns every episode's parameters."""
    rng = np.random.RandomState(4)
    d, B, E, T = 15, 300, 5, 6
    batches = [np.float32(rng.dirichlet(np.ones(d), size=B)) for _ in range(E)]
    kw = dict(T=T, lr_critic=0.1, lr_actor=0.01, seed=9)

    def fresh(w0=None):
        ac = mfg_ac2.actor_critic(theta=8.0, shift=0.16, alpha_scale=12000, d=d, mat_pi0=start_states, seed=9)
        if w0 is not None:
            ac.w = w0.copy()
        return ac
    a = fresh()
    w0 = np.asarray(a.w, dtype=np.float64).copy()
    thetas, ws = [], []
    for e in range(E):
        a.train_batch(batches[e], num_episodes=1, first_episode=e, **kw)
        thetas.append(a.theta)
        ws.append(np.asarray(a.w).reshape(-1).copy())
    pinned = [torch.from_numpy(x).pin_memory() for x in batches]
    res = fresh(w0).train_batch(pinned, num_episodes=E, first_episode=0, history=True, **kw)
    np.testing.assert_allclose(res["theta_history"], thetas, rtol=1e-14)
    np.testing.assert_allclose(res["w_history"], np.stack(ws), rtol=1e-14)
    r2 = fresh(w0).train_batch(lambda e: batches[e], num_episodes=E, first_episode=0, history=True, **kw)
    np.testing.assert_allclose(r2["theta_history"], thetas, rtol=1e-14)
    np.testing.assert_allclose(r2["w_history"], np.stack(ws), rtol=1e-14)


@pytest.mark.parametrize("d", [15, 21])
def test_per_step_update_fused_launch_equals_three_launch_chain(start_states, d):
    """update="per_step": dmfg_ac_step (sampling + TD sums + batch-mean update of theta, w in ONE launch per transition;
    the last CTA reduces the per-CTA partials in CTA order) against rollout(T=1) -> reduce -> apply_update: the same
    parameters bit for bit, at B = 1 (the reference's own semantics, mfg_ac2.py:497-522) and at a ragged batch."""
    rng = np.random.RandomState(d)
    mat = O.synthetic_start_states(n_rows=21, n_cols=30, d=d, seed=4)
    for B in (1, 777):
        pi0 = np.float32(rng.dirichlet(np.ones(d), size=B))
        outs = []
        for fuse in (True, False):
            ac = mfg_ac2.actor_critic(theta=8.0, shift=0.16, alpha_scale=12000, d=d, mat_pi0=mat, dtype="float32", seed=21)
            ac.w = np.linspace(0.1, 0.9, O.num_features(d)).reshape(-1, 1)
            res = ac.train_batch(pi0, num_episodes=2, T=5, lr_critic=0.1, lr_actor=0.01, update="per_step", fuse_step=fuse)
            outs.append((ac.theta, ac.w.ravel().copy(), res["mean_reward"]))
        (t1, w1, m1), (t0, w0, m0) = outs
        assert t1 == t0 and np.array_equal(w1, w0)
        np.testing.assert_allclose(m1, m0, rtol=1e-13)
        assert t1 != 8.0


def test_repeated_train_batch_calls_never_replay_noise(start_states):
    """A loop of train_batch calls with default arguments walks ON through the Philox stream (persistent episode counter):
    two 1-episode calls equal one 2-episode call (with constant step sizes), and the second call's noise is new."""
    rng = np.random.RandomState(5)
    pi0 = np.float32(rng.dirichlet(np.ones(15), size=256))
    kw = dict(T=6, constant=1, lr_critic=0.0, lr_actor=0.0)                # frozen parameters: only the noise differs

    def fresh():
        ac = mfg_ac2.actor_critic(theta=8.0, shift=0.16, alpha_scale=12000, d=15, mat_pi0=start_states, seed=11)
        ac.w = np.full((136, 1), 0.5)
        return ac
    a = fresh()
    r1 = a.train_batch(pi0, num_episodes=1, **kw)["mean_reward"][0]
    r2 = a.train_batch(pi0, num_episodes=1, **kw)["mean_reward"][0]
    both = fresh().train_batch(pi0, num_episodes=2, **kw)["mean_reward"]
    assert r1 != r2
    np.testing.assert_allclose([r1, r2], both, rtol=1e-14)
    # the single-population helpers draw from their own Philox populations (HELPER_POP_OFFSET): sample_action does not
    # return the variates train() consumes at (learner 0, step 0)
    b = fresh()
    pi = start_states[0, :15] / start_states[0, :15].sum()
    P_helper = b.sample_action(pi)
    P_train_pos = mfg_ac2.engine.rollout(b._dev(pi.reshape(1, 15)), 8.0, 0.16, 12000, 1, reward="none", seed=11,
                                         pop_offset=0, step_offset=0, outputs=("actions",))["actions"][0, 0].double().cpu().numpy()
    assert np.abs(P_train_pos - P_helper).max() > 1e-6
    P_own = mfg_ac2.engine.rollout(b._dev(pi.reshape(1, 15)), 8.0, 0.16, 12000, 1, reward="none", seed=11,
                                   pop_offset=mfg_ac2.HELPER_POP_OFFSET, step_offset=0,
                                   outputs=("actions",))["actions"][0, 0].double().cpu().numpy()
    np.testing.assert_array_equal(P_own, P_helper)


def test_train_batch_learns(start_states):
    """The batched per-episode actor-critic climbs the reward it is given: with constant step sizes the mean reward of
    8192 populations rises by two orders of magnitude and theta settles (a policy-gradient sign or TD-error error
    would show up here, not in the step-by-step parity tests)."""
    rng = np.random.RandomState(0)
    ac = mfg_ac2.actor_critic(theta=6.0, shift=0.16, alpha_scale=12000, d=15, mat_pi0=start_states, dtype="float32", seed=1)
    pi0 = np.float32(rng.dirichlet(np.ones(15), size=8192))
    res = ac.train_batch(pi0, num_episodes=60, T=15, constant=1, lr_critic=0.1, lr_actor=1.0, history=True)
    mr, th = res["mean_reward"], res["theta_history"]
    assert np.isfinite(mr).all() and np.isfinite(th).all()
    assert mr[-10:].mean() > 50 * abs(mr[0]) and mr[-10:].mean() > mr[:5].mean()
    assert th[-1] > 6.5 and np.ptp(th[-20:]) < 0.1
