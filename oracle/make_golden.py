#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs the read-only
checkout at /root/reference):

    python oracle/make_golden.py

The reference modules ``mfg_ac2``, ``mfg_ac`` and ``mfg_synthetic`` only need
NumPy/SciPy and are imported unmodified.  Nothing is copied from them: the
script calls their public methods on fixed inputs and stores inputs + outputs.
``ac_irl.py`` / ``networks.py`` need TensorFlow 1.x and cannot be imported, so
there are no reward-net fixtures (parity unpinned at that boundary).

Fixtures
--------
kat_small.npz     the fixed small inputs of test.py / test2.py with the outputs
                  the reference computes for them
train_trace.npz   config 1: mfg_ac2.actor_critic.train, d=15, 3 episodes, every
                  draw (start row, Gamma variates) and every per-step result
traj_d15.npz      generate_trajectory (rollout only) with its recorded draws
forward_d47.npz   test2.test_forward's d=47 start state, 3 transitions
eval_metrics.npz  actor_critic.evaluate (L1 / Jensen-Shannon against empirical days) with its draws
synthetic_check.npz  mfg_synthetic.evaluate_synthetic / evaluate_synthetic_JSD on 3 start rows: the sampled
                  actions and the (mean, std) both functions return
"""
import importlib
import os
import sys
import tempfile
import warnings

import numpy as np

REF = os.environ.get("DMFG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))

from oracle.mfg_oracle import synthetic_start_states  # noqa: E402


def import_reference(name):
    """The reference promotes warnings to errors at import (mfg_ac2.py:21), which
    turns a SyntaxWarning in its own docstring into an ImportError on 3.12."""
    saved = warnings.filters[:]
    warnings.resetwarnings()
    warnings.simplefilter("ignore")
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module(name)
    finally:
        sys.path.remove(REF)
        warnings.filters[:] = saved
    return mod


def write_start_states(dirname, rows20):
    """trend_distribution_day<k>.csv: first line = start row, space separated %.3e
    (the format parsed at mfg_ac2.py:191-198)."""
    os.makedirs(dirname, exist_ok=True)
    for k, row in enumerate(rows20, start=1):
        with open(os.path.join(dirname, "trend_distribution_day%d.csv" % k), "w") as f:
            f.write(" ".join("%.3e" % v for v in row) + "\n")


class Recorder:
    """Records every np.random.gamma / randint call made by the reference."""

    def __init__(self):
        self.gamma_calls, self.randint_calls = [], []
        self._gamma, self._randint = np.random.gamma, np.random.randint

    def __enter__(self):
        def gamma(shape, scale=1.0, size=None):
            y = self._gamma(shape=shape, scale=scale, size=size)
            self.gamma_calls.append(np.array(y, copy=True))
            return y

        def randint(*a, **k):
            v = self._randint(*a, **k)
            self.randint_calls.append(int(v))
            return v

        np.random.gamma, np.random.randint = gamma, randint
        return self

    def __exit__(self, *exc):
        np.random.gamma, np.random.randint = self._gamma, self._randint


def kat_small(mfg_ac2, mfg_ac, mfg_synthetic):
    out = {}
    P3 = np.array([[1, 3, 3], [4, 5, 6], [7, 8, 9]], dtype=float)      # test2.py:49
    pi3 = np.array([0.1, 0.2, 0.7])                                    # test2.py:51
    ac3 = mfg_ac2.actor_critic(d=3)
    out["P3"], out["pi3"] = P3, pi3
    out["reward3"] = np.asarray(ac3.calc_reward(P3.copy(), pi3, 3)).reshape(())
    out["cost3_ac1"] = np.asarray(mfg_ac.actor_critic.calc_cost(None, P3.copy(), pi3, 3)).reshape(())
    ac3.w = np.ones(10)                                                # test2.py:77
    out["value3_w1"] = np.asarray(ac3.calc_value(pi3)).reshape(())
    out["features_235"] = ac3.calc_features(np.array([2.0, 3.0, 5.0]))
    syn = mfg_synthetic.actor_critic.calc_reward(None, P3 / P3.sum(1, keepdims=True), pi3, 3)
    out["P3_rowstochastic"] = P3 / P3.sum(1, keepdims=True)
    out["reward3_synthetic"] = np.asarray(syn).reshape(())

    # test2.py:6,16,105-121  d=4, theta=10, shift=0.4
    ac4 = mfg_ac2.actor_critic(theta=10, shift=0.4, d=4)
    pi4 = np.array([0.7, 0.09, 0.01, 0.2])
    np.random.seed(0)
    with Recorder() as rec:
        P4 = ac4.sample_action(pi4)
    out["d4_theta"], out["d4_shift"], out["d4_alpha_scale"] = 10.0, 0.4, float(ac4.alpha_scale)
    out["d4_pi"] = pi4
    out["d4_y"] = np.stack(rec.gamma_calls)
    out["d4_alpha"], out["d4_alpha_deriv"] = ac4.mat_alpha.copy(), ac4.mat_alpha_deriv.copy()
    out["d4_P"] = P4.copy()
    out["d4_pi_next"] = P4.T.dot(pi4)
    out["d4_reward"] = np.asarray(ac4.calc_reward(P4, pi4, 4)).reshape(())
    out["d4_grad_basic"] = np.asarray(ac4.calc_gradient_basic(P4.copy(), pi4)).reshape(())
    out["d4_grad_loop"] = np.asarray(ac4.calc_gradient(P4.copy(), pi4)).reshape(())
    out["d4_grad_vectorized"] = np.asarray(ac4.calc_gradient_vectorized(P4.copy(), pi4)).reshape(())
    out["d4_features"] = ac4.calc_features(pi4)
    out["d4_reward_synthetic"] = np.asarray(
        mfg_synthetic.actor_critic.calc_reward(None, P4.copy(), pi4, 4)).reshape(())
    np.savez(os.path.join(OUT, "kat_small.npz"), **out)
    print("kat_small: reward3=%r value3=%r d4_grad=%r" % (
        float(out["reward3"]), float(out["value3_w1"]), float(out["d4_grad_vectorized"])))


def train_trace(mfg_ac2, d=15, episodes=3):
    """Config 1: the reference's train() with every intermediate."""
    theta0, shift, alpha_scale = 8.86349, 0.16, 12000
    np.random.seed(0)
    ac = mfg_ac2.actor_critic(theta=theta0, shift=shift, alpha_scale=alpha_scale, d=d)
    w0 = ac.w.copy().reshape(-1)                      # U[0,1) from the seeded global stream
    steps = []

    orig_sample, orig_reward, orig_grad = ac.sample_action, ac.calc_reward, ac.calc_gradient_vectorized

    def sample_action(pi):
        P = orig_sample(pi)
        steps.append(dict(pi=pi.copy(), P=P.copy(), theta_before=float(ac.theta),
                          w_before=ac.w.copy().reshape(-1),
                          alpha=ac.mat_alpha.copy(), alpha_deriv=ac.mat_alpha_deriv.copy()))
        return P

    def calc_reward(P, pi, dd):
        r = orig_reward(P, pi, dd)
        steps[-1]["reward"] = float(np.asarray(r).reshape(()))
        return r

    def calc_gradient_vectorized(P, pi):
        g = orig_grad(P, pi)
        steps[-1]["grad"] = float(g)
        steps[-1]["w_after"] = ac.w.copy().reshape(-1)     # critic already updated (mfg_ac2.py:511-518)
        return g

    ac.sample_action, ac.calc_reward, ac.calc_gradient_vectorized = \
        sample_action, calc_reward, calc_gradient_vectorized
    with Recorder() as rec:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            ac.train(num_episodes=episodes, gamma=1, constant=0, lr_critic=0.1, lr_actor=0.1)
    T = 15
    n = episodes * T
    assert len(steps) == n and len(rec.gamma_calls) == n * d and len(rec.randint_calls) == episodes
    y = np.stack(rec.gamma_calls).reshape(episodes, T, d, d)
    theta_after = [s["theta_before"] for s in steps[1:]] + [float(ac.theta)]
    delta = []
    for s in steps:       # delta exactly as mfg_ac2.py:505, from the reference's own features
        pn = s["P"].T.dot(s["pi"])
        s["pi_next"] = pn
        delta.append(float(s["reward"] + ac.calc_features(pn).dot(s["w_before"])
                           - ac.calc_features(s["pi"]).dot(s["w_before"])))
    np.savez(
        os.path.join(OUT, "train_trace.npz"),
        d=d, T=T, episodes=episodes, theta0=theta0, shift=shift, alpha_scale=float(alpha_scale),
        lr_critic=0.1, lr_actor=0.1, gamma=1.0, mat_pi0=ac.mat_pi0, w0=w0,
        start_rows=np.array(rec.randint_calls), y=y,
        pi=np.stack([s["pi"] for s in steps]).reshape(episodes, T, d),
        P=np.stack([s["P"] for s in steps]).reshape(episodes, T, d, d),
        alpha=np.stack([s["alpha"] for s in steps]).reshape(episodes, T, d, d),
        alpha_deriv=np.stack([s["alpha_deriv"] for s in steps]).reshape(episodes, T, d, d),
        pi_next=np.stack([s["pi_next"] for s in steps]).reshape(episodes, T, d),
        reward=np.array([s["reward"] for s in steps]).reshape(episodes, T),
        grad=np.array([s["grad"] for s in steps]).reshape(episodes, T),
        delta=np.array(delta).reshape(episodes, T),
        theta_after=np.array(theta_after).reshape(episodes, T),
        w_after=np.stack([s["w_after"] for s in steps]).reshape(episodes, T, -1),
        theta_final=float(ac.theta), w_final=ac.w.copy().reshape(-1),
    )
    print("train_trace: theta %.6f -> %.6f after %d steps" % (theta0, ac.theta, n))


def traj_d15(mfg_ac2, d=15):
    np.random.seed(7)
    ac = mfg_ac2.actor_critic(theta=8.86349, shift=0.16, alpha_scale=12000, d=d)
    pi0 = ac.mat_pi0[3].copy()
    with Recorder() as rec:
        traj = ac.generate_trajectory(pi0, 16)
    np.savez(os.path.join(OUT, "traj_d15.npz"), d=d, theta=8.86349, shift=0.16, alpha_scale=12000.0,
             pi0=pi0, y=np.stack(rec.gamma_calls).reshape(15, d, d), trajectory=traj)
    print("traj_d15: final state sum %.6f" % traj[-1].sum())


def forward_d47(mfg_ac2):
    """test2.py:203-224 (start state); theta/shift/alpha_scale are caller arguments there."""
    d = 47
    ac = mfg_ac2.actor_critic(theta=8.86349, shift=0.16, alpha_scale=12000, d=4)
    ac.d = d
    pi = np.zeros(d)
    pi[0] = 0.9
    pi[1:] = np.ones(d - 1) * 0.1 / (d - 1)
    np.random.seed(11)
    pis, Ps, rewards, grads = [pi.copy()], [], [], []
    with Recorder() as rec:
        for _ in range(3):
            P = ac.sample_action(pi)
            rewards.append(float(np.asarray(ac.calc_reward(P, pi, d)).reshape(())))
            grads.append(float(ac.calc_gradient_vectorized(P.copy(), pi)))
            pi = P.T.dot(pi)
            Ps.append(P.copy())
            pis.append(pi.copy())
    np.savez(os.path.join(OUT, "forward_d47.npz"), d=d, theta=8.86349, shift=0.16, alpha_scale=12000.0,
             y=np.stack(rec.gamma_calls).reshape(3, d, d), states=np.stack(pis), actions=np.stack(Ps),
             reward=np.array(rewards), grad=np.array(grads))
    print("forward_d47: rewards", rewards)


def eval_metrics(mfg_ac2, tmp, d=15, n_files=4):
    """mfg_ac2.actor_critic.evaluate (mfg_ac2.py:595-670) on synthetic test days: the empirical
    trajectories, the recorded Gamma draws, the four returned means and the reference's own JSD values."""
    rng = np.random.RandomState(21)
    indir = "test_golden"
    os.makedirs(os.path.join(tmp, indir))
    emp = rng.dirichlet(np.ones(20), size=(n_files, 16))
    emp[1, 5, 3] = 0.0                                        # exercises the zero -> 1e-100 replacement
    for k in range(n_files):
        np.savetxt(os.path.join(tmp, indir, "trend_distribution_day%d.csv" % (22 + k)), emp[k], fmt="%.6e",
                   delimiter=" ")
    emp = np.stack([np.loadtxt(os.path.join(tmp, indir, "trend_distribution_day%d.csv" % (22 + k)))
                    for k in range(n_files)])[:, :, :d]
    ac = mfg_ac2.actor_critic(theta=8.0, shift=0.16, alpha_scale=12000, d=d)
    real_listdir = os.listdir
    os.listdir = lambda path: sorted(real_listdir(path))      # the reference iterates in directory order
    np.random.seed(13)
    try:
        with Recorder() as rec:
            res = ac.evaluate(theta=8.0, shift=0.16, alpha_scale=12000, d=d, episode_length=16, indir=indir,
                              outfile=os.path.join(tmp, "eval.csv"), write_header=1)
    finally:
        os.listdir = real_listdir
    y = np.stack(rec.gamma_calls).reshape(n_files, 15, d, d)
    P = np.array([0.2, 0.0, 0.5, 0.3])
    Q = np.array([0.1, 0.4, 0.0, 0.5])
    np.savez(os.path.join(OUT, "eval_metrics.npz"), d=d, theta=8.0, shift=0.16, alpha_scale=12000.0,
             empirical=emp, y=y, result=np.array(res), csv=np.array(open(os.path.join(tmp, "eval.csv")).read()),
             jsd_P=P, jsd_Q=Q, jsd_value=mfg_ac2.actor_critic.JSD(None, P.copy(), Q.copy()))
    print("eval_metrics:", res)


def synthetic_check(mfg_synthetic, d=15, days=3):
    """mfg_synthetic.actor_critic.evaluate_synthetic / evaluate_synthetic_JSD (mfg_synthetic.py:741-899): the
    actions both calls sample (captured from generate_trajectory) and the (mean, std) they return."""
    ac = mfg_synthetic.actor_critic(theta=2.6, shift=0.5, alpha_scale=1e4, d=d)
    # this variant's constructor does not read the start rows (init_pi0 wants *_reordered.csv files, :181)
    ac.mat_pi0 = synthetic_start_states(n_rows=21, n_cols=20, d=d, seed=0)
    captured = []
    real = ac.generate_trajectory

    def capture(pi0, total_hours):
        traj, acts = real(pi0, total_hours)
        captured.append(np.array(acts, copy=True))
        return traj, acts
    ac.generate_trajectory = capture
    np.random.seed(17)
    l1 = ac.evaluate_synthetic(day_first=1, day_last=days)
    a_l1 = np.stack(captured)
    captured.clear()
    js = ac.evaluate_synthetic_JSD(day_first=1, day_last=days)
    a_js = np.stack(captured)
    np.savez(os.path.join(OUT, "synthetic_check.npz"), d=d, theta=2.6, shift=0.5, alpha_scale=1e4,
             mat_pi0=ac.mat_pi0[:days], actions_l1=a_l1, l1_mean_std=np.array(l1), actions_jsd=a_js,
             jsd_mean_std=np.array(js))
    print("synthetic_check: l1", l1, "jsd", js)


def main():
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        rows20 = synthetic_start_states(n_rows=21, n_cols=20, d=20, seed=0)
        write_start_states(os.path.join(tmp, "train_normalized_round2"), rows20)
        cwd = os.getcwd()
        os.chdir(tmp)            # the reference reads ./train_normalized_round2 from the CWD (mfg_ac2.py:39)
        try:
            mfg_ac2 = import_reference("mfg_ac2")
            mfg_ac = import_reference("mfg_ac")
            mfg_synthetic = import_reference("mfg_synthetic")
            warnings.resetwarnings()
            warnings.simplefilter("ignore")
            kat_small(mfg_ac2, mfg_ac, mfg_synthetic)
            train_trace(mfg_ac2)
            traj_d15(mfg_ac2)
            forward_d47(mfg_ac2)
            eval_metrics(mfg_ac2, tmp)
            synthetic_check(mfg_synthetic)
        finally:
            os.chdir(cwd)


if __name__ == "__main__":
    main()
