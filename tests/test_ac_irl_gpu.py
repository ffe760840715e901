"""AC_IRL drop-in (ac_irl.py:31-954) on the GPU against the oracle: constructor surface, sess.run shim,
train() with the reward net in the loop, update_reward (loss, gradient, Adam), reward_iteration, outerloop.
"""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mfg_oracle as O
from oracle import rnet_oracle as R

D, N3, N4, T = 15, 8, 4, 15


@pytest.fixture(scope="module")
def data():
    rng = np.random.RandomState(0)
    mat = O.synthetic_start_states(n_rows=21, n_cols=20, d=D, seed=0)
    demos = []
    for j in range(8):                                   # synthetic "measured" trajectories at theta* = 8.06
        pi = mat[rng.randint(21)].copy()
        traj = []
        for t in range(T):
            alpha, _ = O.policy_alpha(pi, 8.06, 0.0)
            P = O.normalise_gamma(rng.gamma(alpha * 1e4))
            traj.append((pi, P))
            pi = O.mean_field_step(P, pi)
        demos.append(traj)
    return mat, demos


def make(data, reg="none", **kw):
    from discrete_mean_field_game_b200.ac_irl import AC_IRL
    mat, demos = data
    np.random.seed(1)
    return AC_IRL(theta=6.5, d=D, reg=reg, n_fc3=N3, n_fc4=N4, mat_pi0=mat, demonstrations=demos,
                  seed=5, net_seed=2, **kw)


def test_constructor_surface_and_session_shim(data):
    ac = make(data)
    assert (ac.theta, ac.theta_initial, ac.shift, ac.alpha_scale, ac.d) == (6.5, 6.5, 0, 1e4, D)
    assert ac.lr_reward == 1e-4 and ac.num_policies == 10 and ac.c == 2e11
    assert ac.w.shape == (136, 1) and ac.num_start_samples == 21
    assert ac.num_demo_samples == 5 and ac.num_gen_samples == 5 and ac.num_sampled_trajectories == 5
    assert ac.list_policies == [6.5] * 10 and ac.list_generated == []
    assert len(ac.list_eval_demo_transitions) == 8 * T
    assert ac.reward_params.count == 3755
    # test_acirl.py:60-70: sess.run(reward_gen, {gen_states: [...], gen_actions: [...]})
    s = [p[0] for p in data[1][0]]
    a = [p[1] for p in data[1][0]]
    r = ac.sess.run(ac.reward_gen, feed_dict={ac.gen_states: s, ac.gen_actions: a})
    assert r.shape == (T, 1)
    ref = R.forward(ac.reward_params.flat.cpu().numpy(), np.float32(s), np.float32(a), N3, N4)
    np.testing.assert_allclose(r[:, 0], ref, rtol=1e-5, atol=1e-6)
    rd, rg = ac.sess.run([ac.reward_demo, ac.reward_gen], feed_dict={ac.demo_states: s, ac.demo_actions: a,
                                                                    ac.gen_states: s[:3], ac.gen_actions: a[:3]})
    np.testing.assert_allclose(rd, r, rtol=1e-6)
    assert rg.shape == (3, 1)


def test_generate_trajectories_structure(data):
    ac = make(data)
    trajs = ac.generate_trajectories(6)
    assert len(trajs) == 6 and all(len(t) == T for t in trajs)
    s0, a0 = trajs[2][0]
    s1, _ = trajs[2][1]
    assert s0.shape == (D,) and a0.shape == (D, D)
    np.testing.assert_allclose(a0.sum(1), 1.0, atol=1e-6)
    np.testing.assert_allclose(a0.T.dot(s0), s1, atol=3e-7)             # test_acirl.py:43-47
    assert any(np.allclose(s0, row, atol=1e-7) for row in data[0])     # started from a training row


def test_train_matches_serial_oracle_with_reward_net(data):
    """3 episodes of AC_IRL.train with injected draws vs. the oracle's serial learner whose reward is
    the oracle reward net on the same float32 parameters (ac_irl.py:634-732)."""
    ac = make(data)
    rng = np.random.RandomState(3)
    E = 3
    rows = rng.randint(21, size=E)
    params = ac.reward_params.flat.cpu().numpy().astype(np.float64) + 0.2 * rng.randn(3755)   # non-trivial rewards
    ac.reward_params.load_flat(params)
    params = ac.reward_params.flat.cpu().numpy()
    w0 = ac.w.copy()
    # draw the Gamma variates along the ORACLE trajectory (float32-rounded so both sides see the same y)
    ys = np.zeros((E, T, D, D), np.float32)

    class Noise:
        e, t = -1, 0

        def start_index(self, n):
            Noise.e += 1
            Noise.t = 0
            return int(rows[Noise.e])

        def gamma_rows(self, shape):
            y = np.float32(rng.gamma(shape))
            ys[Noise.e, Noise.t] = y
            Noise.t += 1
            return y.astype(np.float64)

    def reward_fn(pi, P):
        return R.forward(params, np.float32(pi)[None], np.float32(P)[None], N3, N4)[0]

    th, w, info = O.train_serial(data[0], 6.5, w0, 0.0, 1e4, E, lr_critic=0.1, lr_actor=0.001, flavour="ac_irl",
                                 reward_fn=reward_fn, noise=Noise(), num_steps=T)
    ac.train(max_episodes=E, stop_criteria=-1, lr_critic=0.1, lr_actor=0.001, start_rows=rows, noise_y=ys,
             verbose=False)
    np.testing.assert_allclose(ac.theta, th, rtol=1e-5)
    np.testing.assert_allclose(ac.w.ravel(), w, rtol=1e-5, atol=1e-6)
    assert ac.list_policies[-1] == ac.theta and len(ac.list_policies) == 10


@pytest.mark.parametrize("reg", ["none", "l1l2"])
def test_update_reward_matches_oracle(data, reg):
    ac = make(data, reg=reg)
    ac.list_generated = ac.generate_trajectories(10)
    p0 = ac.reward_params.flat.cpu().numpy().astype(np.float64)
    m = np.zeros_like(p0)
    v = np.zeros_like(p0)
    for step in range(1, 4):
        random.seed(100 + step)
        ac.update_reward()
        random.seed(100 + step)
        demo = random.sample(ac.list_demonstrations, 5)
        gen = random.sample(ac.list_generated, 5)
        ds = np.float32([p[0] for tr in demo for p in tr]); da = np.float32([p[1] for tr in demo for p in tr])
        gs = np.float32([p[0] for tr in gen for p in tr]); ga = np.float32([p[1] for tr in gen for p in tr])
        loss, (first, second), grad = R.loss_and_grad(p0, ds, da, gs, ga, N3, N4, 5, T, reg=reg)
        np.testing.assert_allclose([ac.loss_val, ac.first_term_val, ac.second_term_val], [loss, first, second],
                                   rtol=1e-5, atol=1e-6)
        g_dev = ac._last_grad.cpu().numpy()
        g_data = grad - (R.reg_grad(p0, D, N3, N4) if reg == "l1l2" else 0)
        assert np.abs(g_dev - g_data).max() <= 1e-5 * np.abs(g_data).max() + 1e-6
        p0, m, v = R.adam_tf(p0, m, v, grad, step, 1e-4)
        # Adam normalises the step to ~lr, so a tiny gradient error can flip a coordinate early on:
        # compare the parameters at the scale of one step
        assert np.abs(ac.reward_params.flat.cpu().numpy() - p0).max() <= 2e-5
        p0 = ac.reward_params.flat.cpu().numpy().astype(np.float64)       # re-sync (float32 storage)
        m = ac.reward_params.m.cpu().numpy().astype(np.float64)
        v = ac.reward_params.v.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("reg", ["none", "dropout_l1l2"])
def test_one_pass_reward_update_equals_the_three_launch_chain(data, reg):
    """update_reward_batch with the generated half in one launch (weight exp(R_j), 1/Z afterwards) against the
    forward -> loss -> backward chain on the same batch, same Philox dropout masks: loss terms, gradient and the
    parameters after the Adam step agree to float accuracy."""
    outs = []
    for one_pass in (True, False):
        irl = make(data, reg=reg)
        irl.one_pass_reward_update = one_pass
        ds, da = irl.generate_batch(37, theta=8.06)
        gs, ga = irl.generate_batch(53)
        T = 15
        loss = irl.update_reward_batch(ds[:T].reshape(-1, 15), da.reshape(-1, 15, 15), gs[:T].reshape(-1, 15),
                                       ga.reshape(-1, 15, 15), 37, "time_major", group=False)
        outs.append((loss.cpu().numpy(), irl._last_grad.cpu().numpy(), irl.reward_params.flat.cpu().numpy()))
    (l1, g1, p1), (l0, g0, p0) = outs
    np.testing.assert_allclose(l1[:3], l0[:3], rtol=1e-5, atol=3e-5)
    assert np.abs(g1 - g0).max() <= 1e-5 * np.abs(g0).max() + 1e-7
    np.testing.assert_allclose(p1, p0, rtol=0, atol=2.1e-4)      # one Adam step moves every weight by <= lr = 1e-4
    assert np.mean(np.abs(p1 - p0) <= 1e-6) > 0.99                # (sign flips of ~0 gradients aside)


def test_irl_step_batch_equals_train_batch_then_update_reward_batch(data):
    """irl_step_batch (one reward-net pass over the generated record serves the forward solve AND the reward update)
    against train_batch(keep_record) followed by update_reward_batch: theta, w, loss and reward parameters."""
    rng = np.random.RandomState(4)
    pi0 = np.float32(rng.dirichlet(np.ones(D), size=45))
    outs = []
    for fused in (True, False):
        irl = make(data, reg="l1l2")
        ds, da = irl.generate_batch(21, theta=8.06)
        ds, da = ds[:T].reshape(-1, D), da.reshape(-1, D, D)
        for ep in (1, 2):
            if fused:
                res = irl.irl_step_batch(pi0, ds, da, 21, episode=ep, lr_critic=0.1, lr_actor=0.01, group=False)
                loss = res["loss"]
            else:
                res = irl.train_batch(pi0, 1, lr_critic=0.1, lr_actor=0.01, first_episode=ep, keep_record=True, group=False)
                loss = irl.update_reward_batch(ds, da, res["states"][:T].reshape(-1, D), res["actions"].reshape(-1, D, D),
                                               21, "time_major", group=False)
        outs.append((irl.theta, irl.w.ravel().copy(), loss.cpu().numpy(), irl.reward_params.flat.cpu().numpy(),
                     irl._last_grad.cpu().numpy()))
    (t1, w1, l1, p1, g1), (t0, w0, l0, p0, g0) = outs
    np.testing.assert_allclose(t1, t0, rtol=1e-9)
    np.testing.assert_allclose(w1, w0, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(l1[:3], l0[:3], rtol=1e-5, atol=3e-5)
    assert np.abs(g1 - g0).max() <= 1e-5 * np.abs(g0).max() + 1e-7
    assert np.mean(np.abs(p1 - p0) <= 1e-6) > 0.99 and np.abs(p1 - p0).max() <= 4.2e-4     # two Adam steps of 1e-4


def test_data_parallel_reward_step_is_rank_count_invariant(data):
    """The data-parallel reward step (ac_irl.py:390-406 over the union of all ranks' trajectories): every "rank"
    contributes RAW sums (demonstration gradient for dL/dr = -1, unnormalised generated gradient sum_j e^{R_j} dR_j,
    Z, sum r_demo, counts); 1/N_demo and 1/Z are applied after the sum.  Two and three virtual ranks on one GPU must
    give the single-batch gradient, loss and Adam step (1e-6), whatever the (uneven) sharding."""
    from discrete_mean_field_game_b200 import engine
    irl = make(data, reg="none")
    ds, da = irl.generate_batch(37, theta=8.06)
    gs, ga = irl.generate_batch(53)
    P = irl.reward_params.flat.numel()

    def shard(x, lo, hi, tail):
        return x[:T, lo:hi].reshape((-1,) + tail).contiguous()

    ref_irl = make(data, reg="none")
    loss_ref = ref_irl.update_reward_batch(shard(ds, 0, 37, (D,)), shard(da, 0, 37, (D, D)), shard(gs, 0, 53, (D,)),
                                           shard(ga, 0, 53, (D, D)), 37, "time_major", group=False).cpu().numpy()
    g_ref = ref_irl._last_grad.cpu().numpy()
    for cuts_d, cuts_g in (((0, 20, 37), (0, 11, 53)), ((0, 1, 30, 37), (0, 25, 26, 53))):
        total = torch.zeros(2 * P + 4, dtype=torch.float64, device=irl.device)
        for k in range(len(cuts_d) - 1):
            terms, _ = irl._dp_reward_terms(shard(ds, cuts_d[k], cuts_d[k + 1], (D,)), shard(da, cuts_d[k], cuts_d[k + 1], (D, D)),
                                            shard(gs, cuts_g[k], cuts_g[k + 1], (D,)), shard(ga, cuts_g[k], cuts_g[k + 1], (D, D)),
                                            cuts_d[k + 1] - cuts_d[k], "time_major")
            total += terms
        grad, loss = engine.irl_dp_finalize(total, P)
        assert float(total[2 * P + 2]) == 37 and float(total[2 * P + 3]) == 53
        np.testing.assert_allclose(loss.cpu().numpy()[:3], loss_ref[:3], rtol=1e-6, atol=1e-6)
        err = np.abs(grad.cpu().numpy() - g_ref).max()
        assert err <= 1e-5 * np.abs(g_ref).max() + 1e-8, (err, np.abs(g_ref).max())     # float32 summation order only
    # the forced single-rank form of the public call takes the same path and lands on the same parameters
    irl2 = make(data, reg="none")
    irl2.rank_invariant_reward_step = True
    irl2.update_reward_batch(shard(ds, 0, 37, (D,)), shard(da, 0, 37, (D, D)), shard(gs, 0, 53, (D,)),
                             shard(ga, 0, 53, (D, D)), 37, "time_major", group=False)
    p2, p1 = irl2.reward_params.flat.cpu().numpy(), ref_irl.reward_params.flat.cpu().numpy()
    assert np.mean(np.abs(p2 - p1) <= 1e-6) > 0.99 and np.abs(p2 - p1).max() <= 2.1e-4   # Adam: sign flips of ~0 gradients aside


def test_write_all_dumps_every_step_like_the_reference(data, tmp_path, monkeypatch):
    """write_all=1 appends 'Episode k', then per step the state and the sampled P to temp.csv (ac_irl.py:651-676,
    mfg_ac2.py:461-494)."""
    from discrete_mean_field_game_b200.mfg_ac2 import actor_critic
    monkeypatch.chdir(tmp_path)
    ac = actor_critic(theta=8.0, shift=0.1, alpha_scale=1e4, d=D, mat_pi0=data[0], seed=3)
    ac.train(num_episodes=2, lr_critic=0.1, lr_actor=0.01, write_all=1, verbose=False)
    txt = open(tmp_path / "temp.csv").read()
    assert txt.count("Episode") == 2 and txt.count("num_steps = ") == 2 * T and txt.count("Action") == 2 * T
    assert "Episode 0 \n\n" in txt and "num_steps = 15" in txt
    block = txt.split("Action\n")[1].split("num_steps")[0].strip().splitlines()
    rows = np.array([[float(v) for v in ln.split(",")] for ln in block[:D]])
    assert rows.shape == (D, D) and np.allclose(rows.sum(1), 1.0, atol=0.01)
    os.remove(tmp_path / "temp.csv")
    irl = make(data)
    irl.train(max_episodes=2, stop_criteria=-1, write_all=1, verbose=False)
    txt = open(tmp_path / "temp.csv").read()
    assert txt.count("Episode") == 2 and txt.count("distribution") == 2 * T and "Episode 1 \n\n" in txt


def test_update_reward_with_importance_weights(data):
    ac = make(data, use_z=True)
    ac.list_policies = list(np.linspace(6.0, 7.0, 10))
    ac.list_generated = ac.generate_trajectories(10)
    p0 = ac.reward_params.flat.cpu().numpy().astype(np.float64)
    random.seed(7)
    ac.update_reward()
    random.seed(7)
    demo = random.sample(ac.list_demonstrations, 5)
    gen = random.sample(ac.list_generated, 5)
    ds = np.float32([p[0] for tr in demo for p in tr]); da = np.float32([p[1] for tr in demo for p in tr])
    gs = np.float32([p[0] for tr in gen for p in tr]); ga = np.float32([p[1] for tr in gen for p in tr])
    lz = R.log_z(gs.reshape(5, T, D), ga.reshape(5, T, D, D), ac.list_policies, 0.0, 21)
    loss, (first, second), grad = R.loss_and_grad(p0, ds, da, gs, ga, N3, N4, 5, T, log_z=lz)
    np.testing.assert_allclose(ac.second_term_val, second, rtol=1e-5)
    g_dev = ac._last_grad.cpu().numpy()
    assert np.abs(g_dev - grad).max() <= 1e-5 * np.abs(grad).max() + 1e-6


def test_dropout_variant_runs_and_is_stochastic(data):
    ac = make(data, reg="dropout_l1l2")
    s = [p[0] for p in data[1][0]]
    a = [p[1] for p in data[1][0]]
    ac.reward_params.load_flat(ac.reward_params.flat.cpu().numpy() + 0.3)
    r1 = ac.sess.run(ac.reward_gen, feed_dict={ac.gen_states: s, ac.gen_actions: a})
    r2 = ac.sess.run(ac.reward_gen, feed_dict={ac.gen_states: s, ac.gen_actions: a})
    assert not np.array_equal(r1, r2)                   # fresh masks per run, active at inference (quirk C.8)
    ac.list_generated = ac.generate_trajectories(10)
    ac.update_reward()
    assert np.isfinite(ac.loss_val) and ac.loss_val > ac.first_term_val + ac.second_term_val   # + regulariser


def test_outerloop_smoke_and_checkpoint(data, tmp_path):
    ac = make(data)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        theta = ac.outerloop(num_iterations=2, num_gen_from_policy=5, max_reward_iterations=20,
                             max_forward_episodes=4, final_episodes=6, verbose=False)
        assert np.isfinite(theta) and theta == ac.theta
        assert len(ac.list_generated) == 50 and len(ac.list_policies) == 10
        assert ac.list_policies[-1] == ac.theta
        assert os.path.exists("results/reward_training.csv")       # theta.csv only at multiples of 100 episodes
        ck = "log/model_none_8_4.ckpt.npz"
        assert os.path.exists(ck)
        before = ac.reward_params.flat.clone()
        ac.reward_params.initialize(seed=99)
        assert not torch.equal(before, ac.reward_params.flat)
        ac.restore("log/model_none_8_4.ckpt")
        assert torch.equal(before, ac.reward_params.flat)
        lines = open("results/reward_training.csv").read().strip().splitlines()
        assert lines[0] == "reward_demo_avg,reward_gen_avg" and len(lines) >= 3
    finally:
        os.chdir(cwd)


def test_reads_reference_file_formats(data, tmp_path):
    """trend_distribution_day<k>.csv (16 rows) and action_day<k>.txt (15 blocks of 20x20) as
    read_demonstrations / init_pi0 parse them (ac_irl.py:164-200, 443-506)."""
    from discrete_mean_field_game_b200.ac_irl import AC_IRL
    rng = np.random.RandomState(4)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for sd, adir, start, n in (("train_normalized_round2", "actions_2", 1, 3),
                                   ("test_normalized_round2", "actions_test_2", 22, 2)):
            os.makedirs(sd)
            os.makedirs(adir)
            for k in range(start, start + n):
                st = rng.dirichlet(np.ones(20), size=16)
                np.savetxt(os.path.join(sd, "trend_distribution_day%d.csv" % k), st, fmt="%.3e", delimiter=" ")
                with open(os.path.join(adir, "action_day%d.txt" % k), "w") as f:
                    for h in range(15):
                        np.savetxt(f, rng.dirichlet(np.ones(20), size=20), fmt="%.3e", delimiter=" ")
                        f.write("\n")
        ac = AC_IRL(d=15, reg="none", seed=1)
        assert ac.mat_pi0.shape == (3, 15) and ac.mat_pi0_test.shape == (2, 15)
        assert len(ac.list_demonstrations) == 3 and len(ac.list_demonstrations_test) == 2
        s, a = ac.list_demonstrations[1][4]
        assert s.shape == (15,) and a.shape == (15, 15)
        first = np.loadtxt("train_normalized_round2/trend_distribution_day2.csv")[4, :15]
        np.testing.assert_array_equal(s, first)
    finally:
        os.chdir(cwd)


def test_train_batch_with_reward_net_matches_oracle(data):
    """AC_IRL.train_batch: the batched forward solve with r = r_net(pi, P) (rollout + record -> reward net -> TD
    sums -> batch-mean update, cumulative discount of ac_irl.py:691) against the oracle fed the same Gamma draws."""
    import math
    ac = make(data)
    rng = np.random.RandomState(11)
    params = ac.reward_params.flat.cpu().numpy().astype(np.float64) + 0.2 * rng.randn(3755)
    ac.reward_params.load_flat(params)
    params = ac.reward_params.flat.cpu().numpy()
    E, B, gamma = 2, 40, 0.95
    pi0 = np.float32(rng.dirichlet(np.ones(D), size=B))
    theta, w = 6.5, ac.w.ravel().copy()
    ys = np.zeros((E, T, B, D, D), np.float32)
    th_ref, w_ref, mean_r = [], [], []
    for e in range(E):
        episode = 1 + e
        pi = pi0.astype(np.float64)
        for t in range(T):                                   # draws along the oracle's own trajectory
            alpha, _ = O.policy_alpha(pi, theta, 0.0)
            ys[e, t] = np.float32(rng.gamma(alpha * 1e4))
            pi = O.mean_field_step(O.normalise_gamma(ys[e, t].astype(np.float64)), pi)
        first = O.rollout_frozen(pi0.astype(np.float64), theta, 0.0, 1e4, ys[e].astype(np.float64), reward="none")
        r = R.forward(params, np.float32(first["states"][:T].reshape(-1, D)), np.float32(first["actions"].reshape(-1, D, D)),
                      N3, N4).reshape(T, B)
        ref = O.rollout_frozen(pi0.astype(np.float64), theta, 0.0, 1e4, ys[e].astype(np.float64), w=w, gamma=gamma,
                               discount="cumulative", reward=r)
        lr_c = 0.1 / (episode + 1.0)
        lr_a = 0.001 / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
        theta = theta + lr_a / B * ref["G_theta"]
        w = w + lr_c / B * ref["G_w"]
        th_ref.append(theta); w_ref.append(w.copy()); mean_r.append(ref["R"] / B)
    res = ac.train_batch(pi0, num_episodes=E, gamma=gamma, lr_critic=0.1, lr_actor=0.001, noise_y=ys, keep_record=True)
    np.testing.assert_allclose(res["theta"], th_ref[-1], rtol=1e-6)
    np.testing.assert_allclose(ac.w.ravel(), w_ref[-1], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(res["mean_reward"], mean_r, rtol=2e-4, atol=2e-5)
    assert res["states"].shape == (T + 1, B, D) and res["actions"].shape == (T, B, D, D)
    assert ac.list_policies[-1] == ac.theta
    # sampled noise: runs, finite, reproducible
    a1, a2 = make(data), make(data)
    r1 = a1.train_batch(pi0, num_episodes=2)
    r2 = a2.train_batch(pi0, num_episodes=2)
    assert np.isfinite(r1["theta"]) and r1["theta"] == r2["theta"]


def test_train_cuda_graph_replay_equals_eager_launches(data):
    """AC_IRL.train captures the 60-launch chain of an episode as a CUDA graph and replays it with the start state,
    step sizes and Philox position refreshed in device buffers: bit-identical to launching the chain eagerly."""
    a, b = make(data), make(data)
    b.w = a.w.copy()
    kw = dict(max_episodes=6, stop_criteria=-1, lr_critic=0.1, lr_actor=0.01, verbose=False, fused=False)
    a.train(use_graph=True, **kw)
    b.train(use_graph=False, **kw)
    assert a.theta == b.theta and a.theta != 6.5
    np.testing.assert_array_equal(a.w, b.w)
    a.train(use_graph=True, **kw)                     # a second call continues the noise stream in both modes
    b.train(use_graph=False, **kw)
    assert a.theta == b.theta
    np.testing.assert_array_equal(a.w, b.w)


@pytest.mark.parametrize("reg", ["none", "dropout_l1l2"])
def test_train_fused_kernel_equals_host_driven_chain(data, reg, capsys):
    """AC_IRL.train as ONE kernel (dmfg_irl_learners: a CTA per learner, the reward net evaluated inside the loop, Philox
    dropout masks keyed by the transition id) against the host-driven chain of 4 launches per transition on the same
    Philox draws: same start rows, same rewards and TD errors to float32 accuracy, same theta / w after 7 episodes,
    same reports; and the |delta theta| stop criterion fires at the same episode."""
    outs = []
    for fused in (True, False):
        ac = make(data, reg=reg)
        rng = np.random.RandomState(3)
        ac.reward_params.load_flat(ac.reward_params.flat.cpu().numpy() + np.float32(0.2 * rng.randn(3755)))
        ac.w = np.linspace(0.2, 0.8, 136).reshape(-1, 1)
        ac.train(max_episodes=7, stop_criteria=-1, lr_critic=0.1, lr_actor=0.01, consecutive=3, verbose=True, fused=fused,
                 use_graph=False)
        first = (ac.theta, ac.w.ravel().copy(), capsys.readouterr().out)
        ac.train(max_episodes=40, stop_criteria=2e-3, lr_critic=0.1, lr_actor=0.01, verbose=True, fused=fused, use_graph=False)
        outs.append(first + (ac.theta, capsys.readouterr().out.strip().splitlines()[-1]))
    (t1, w1, o1, s1, e1), (t0, w0, o0, s0, e0) = outs
    assert t1 != 6.5
    np.testing.assert_allclose(t1, t0, rtol=2e-5)
    np.testing.assert_allclose(w1, w0, rtol=2e-5, atol=1e-6)
    assert o1.count("Average reward during previous 3 episodes") == 2 == o0.count("Average reward during previous 3 episodes")
    assert e1.split("with theta")[0] == e0.split("with theta")[0] and "Exiting train at episode" in e1   # same stop episode
    np.testing.assert_allclose(s1, s0, rtol=1e-4)


def test_train_fused_many_learners_and_traces(data):
    """dmfg_irl_learners with several learners (a CTA each): learner l equals a single-learner launch with learner_offset l;
    the reward trace equals dmfg_rnet_forward on the recorded transitions is implied by the previous test -- here the
    per-step traces are finite and tanh-bounded."""
    from discrete_mean_field_game_b200 import engine
    ac = make(data)
    dev = ac.device
    mat = torch.as_tensor(np.float32(data[0][:, :D]), device=dev)
    L, E = 5, 4
    th = torch.full((L,), 6.5, dtype=torch.float64, device=dev)
    w = torch.rand((L, 136), dtype=torch.float64, device=dev)
    th1, w1 = th.clone(), w.clone()
    kw = dict(shift=0.0, alpha_scale=1e4, episode0=1, lr_critic=0.1, lr_actor=0.01, seed=9)
    res = engine.irl_learners(th, w, mat, E, T, ac.reward_params.flat, N3, N4, trace=True, **kw)
    assert torch.isfinite(res["theta_trace"]).all() and res["reward_trace"].abs().max() <= 1.0
    for l in (0, 3):
        tl, wl = th1[l:l + 1].clone(), w1[l:l + 1].clone()
        engine.irl_learners(tl, wl, mat, E, T, ac.reward_params.flat, N3, N4, learner_offset=l, **kw)
        assert tl[0] == th[l] and torch.equal(wl[0], w[l])


def test_gridsearch_dropin(data, tmp_path, monkeypatch):
    """gridsearch.py:1-31 with in-memory data and shortened loops: one CSV line per (reg, n_fc3, n_fc4) point,
    test_reward_network (ac_irl.py:1008-1043) returns the three averages."""
    from discrete_mean_field_game_b200 import gridsearch
    monkeypatch.chdir(tmp_path)
    mat, demos = data
    rows = gridsearch.run(outfile=str(tmp_path / "results" / "grid.csv"), list_reg=["none", "dropout_l1l2"],
                          list_nfc3=[8], list_nfc4=[4, 6],
                          ac_kwargs=dict(d=D, mat_pi0=mat, demonstrations=demos, demonstrations_test=demos[:3], seed=5,
                                         net_seed=2),
                          outerloop_kwargs=dict(num_iterations=1, max_reward_iterations=10, max_forward_episodes=4,
                                                final_episodes=4, verbose=False))
    assert len(rows) == 4
    lines = (tmp_path / "results" / "grid.csv").read_text().strip().split("\n")
    assert lines[0].startswith("reg,n_fc3,n_fc4,") and len(lines) == 5
    for reg, n3, n4, tr, te, ge, th in rows:
        assert np.isfinite([tr, te, ge, th]).all() and -1.0 <= tr <= 1.0 and -1.0 <= ge <= 1.0


def test_evaluate_surface_of_ac_irl(data, tmp_path, evalm):
    """ac_irl.py:1445-1589: JSD / generate_trajectory / evaluate exist on AC_IRL with its own defaults and give the
    reference's numbers on the golden evaluation fixture (same code path as mfg_ac2's, pinned there)."""
    ac = make(data)
    out = tmp_path / "validation.csv"
    res = ac.evaluate(theta=float(evalm["theta"]), shift=float(evalm["shift"]), alpha_scale=float(evalm["alpha_scale"]),
                      outfile=str(out), write_header=1, empirical=evalm["empirical"], y=evalm["y"])
    np.testing.assert_allclose(res, evalm["result"], rtol=2e-5)
    assert out.read_text().startswith("theta,shift,alpha_scale,mean_l1_final")
    np.testing.assert_allclose(ac.JSD(evalm["jsd_P"].copy(), evalm["jsd_Q"].copy()), float(evalm["jsd_value"]), rtol=1e-12)
    traj = ac.generate_trajectory(data[0][0], 16)
    assert traj.shape == (16, D)


def test_ac_irl_runs_at_d21_on_20x20_style_data(tmp_path, monkeypatch):
    """AC_IRL at d = 21 / 20 (round 1 returned DMFG_ERR_UNSUPPORTED): the reward net runs on the 32-lane kernels, the reward
    update takes the forward -> loss -> backward chain (the one-pass launch holds a trajectory per CTA at d <= 16 only),
    train() runs host-driven; one outer-loop-style sequence stays finite and moves theta and the reward parameters."""
    from discrete_mean_field_game_b200.ac_irl import AC_IRL
    monkeypatch.chdir(tmp_path)
    for d in (21, 20):
        rng = np.random.RandomState(d)
        mat = rng.dirichlet(np.ones(d), size=12)
        ac = AC_IRL(theta=6.5, d=d, reg="none", mat_pi0=mat, demonstrations=[], seed=3, net_seed=4)
        ds, da = ac.generate_batch(9, theta=8.0)
        gs, ga = ac.generate_batch(11)
        p0 = ac.reward_params.flat.clone()
        loss = ac.update_reward_batch(ds[:15].reshape(-1, d).contiguous(), da.reshape(-1, d, d), gs[:15].reshape(-1, d).contiguous(),
                                      ga.reshape(-1, d, d), 9, "time_major", group=False).cpu().numpy()
        assert np.isfinite(loss).all() and not torch.equal(p0, ac.reward_params.flat)
        # against the oracle: loss and gradient of the same batch
        g_ref = R.loss_and_grad(p0.cpu().numpy(), ds[:15].reshape(-1, d).cpu().numpy(), da.reshape(-1, d, d).cpu().numpy(),
                                gs[:15].permute(1, 0, 2).reshape(-1, d).cpu().numpy(),
                                ga.permute(1, 0, 2, 3).reshape(-1, d, d).cpu().numpy(), 8, 4, 9, 15)
        np.testing.assert_allclose(loss[0], g_ref[0], rtol=1e-4, atol=1e-5)
        g = ac._last_grad.cpu().numpy()
        assert np.abs(g - g_ref[2]).max() <= 3e-5 * np.abs(g_ref[2]).max() + 1e-6
        ac.train(max_episodes=2, stop_criteria=-1, lr_critic=0.1, lr_actor=0.01, verbose=False)
        assert np.isfinite(ac.theta) and ac.theta != 6.5 and np.isfinite(ac.w).all()


@pytest.mark.parametrize("reg", ["none", "dropout_l1l2"])
def test_reward_update_through_one_c_call_equals_the_three_call_chain(data, reg):
    """update_reward on one rank goes through dmfg_irl_reward_step (rnet_backward -> rnet_backward_gen -> adam_tf behind one
    entry point, the reductions / loss terms / Adam step behind the two backward launches in ONE finishing launch): same
    summation order and roundings as the three separate calls -- parameters, Adam moments and the gradient agree bit for
    bit over several updates, with and without the dropout / l1l2 regulariser; the loss terms to the last bits of a double
    (the finishing launch sums per-CTA reward sums, the chain walks r_demo)."""
    import random
    outs = []
    for fused in (True, False, "chain"):
        ac = make(data, reg=reg)
        ac.fused_reward_step = fused
        ac.list_generated = ac.generate_trajectories(12)
        random.seed(7)
        losses = []
        for _ in range(4):
            ac.update_reward()
            losses.append((ac.loss_val, ac.first_term_val, ac.second_term_val))
        p = ac.reward_params
        outs.append((p.flat.clone(), p.m.clone(), p.v.clone(), ac._last_grad.clone(), losses, p.step))
    a, b, c = outs
    assert a[5] == b[5] == c[5] == 4
    for x, y, z in zip(a[:4], b[:4], c[:4]):
        assert torch.equal(x, y) and torch.equal(y, z)
    assert b[4] == c[4]                                   # the same launches in the same order
    np.testing.assert_allclose(np.array(a[4]), np.array(b[4]), rtol=1e-13, atol=1e-14)
    assert not torch.equal(a[0], make(data, reg=reg).reward_params.flat)          # the parameters did move


@pytest.mark.parametrize("reg", ["none", "dropout_l1l2"])
def test_reward_update_on_resident_trajectories_equals_the_stacked_update(data, reg):
    """update_reward with the sampled trajectories resident in a device pool (uploaded once, the minibatch = slot numbers
    by value with the launch) against the same updates on freshly stacked and copied arrays: bit-identical parameters,
    moments, gradient and loss terms -- across pool growth (more trajectories than the first 64 slots) and with
    trajectories appended to list_generated between updates."""
    outs = []
    for resident in (True, False):
        ac = make(data, reg=reg)
        ac.resident_trajectories = resident
        ac.list_generated = ac.generate_trajectories(70)
        random.seed(11)
        losses = []
        for k in range(30):
            ac.update_reward()
            losses.append((ac.loss_val, ac.first_term_val, ac.second_term_val))
            if k == 12:
                ac.list_generated = ac.list_generated + ac.generate_trajectories(30)
        p = ac.reward_params
        outs.append((p.flat.clone(), p.m.clone(), p.v.clone(), ac._last_grad.clone(), losses))
        if resident:
            pool = ac._pool
            assert 64 < pool["used"] <= 108 and pool["states"].shape[0] == 128 * T      # grew once
            assert len(pool["index"]) == pool["used"]
        else:
            assert "_pool" not in ac.__dict__
    a, b = outs
    for x, y in zip(a[:4], b[:4]):
        assert torch.equal(x, y)
    assert a[4] == b[4]


def test_resident_pool_starts_over_when_full(data):
    ac = make(data)
    ac._POOL_LIMIT = 12
    ac.list_generated = ac.generate_trajectories(20)
    ref = make(data)
    ref.resident_trajectories = False
    ref.list_generated = ac.list_generated
    ref.list_demonstrations = ac.list_demonstrations
    for seed in range(8):
        random.seed(seed); ac.update_reward()
        random.seed(seed); ref.update_reward()
        assert ac._pool["used"] <= 12
    assert torch.equal(ac.reward_params.flat, ref.reward_params.flat)
