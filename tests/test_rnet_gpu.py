"""Reward-net forward/backward, IRL loss, TF-Adam and calc_z kernels against the float64 oracle (B200).

float32 kernels vs. float64 oracle on identical float32-rounded inputs; tolerances stated per check.
The oracle itself is checked against torch autograd in tests/test_rnet_oracle.py (TF boundary unpinned).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rnet_oracle as R


@pytest.fixture(scope="module")
def dev():
    from discrete_mean_field_game_b200 import engine
    engine.require_cuda()
    return torch.device("cuda:0")


def f32(x):
    return np.asarray(x, dtype=np.float32)


def make(rng, n, d, n3, n4, scale=0.05):
    p = f32(R.xavier_init(d, n3, n4, rng) + scale * rng.randn(R.param_count(d, n3, n4)))
    s = f32(rng.dirichlet(np.ones(d), size=n))
    a = f32(rng.dirichlet(np.ones(d) * 0.5, size=(n, d)))
    return p, s, a


def T_(x, dev, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=dev)


@pytest.mark.parametrize("d,n3,n4", [(15, 8, 4), (15, 4, 6), (4, 6, 8), (16, 8, 8), (7, 5, 3)])
@pytest.mark.parametrize("dropout", [False, True])
def test_forward_matches_oracle(dev, d, n3, n4, dropout):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(d * 100 + n3)
    n = 1000 + d                                   # not a multiple of the 16-transition tile
    p, s, a = make(rng, n, d, n3, n4)
    m3 = (rng.rand(n, n3) < 0.4) if dropout else None
    m4 = (rng.rand(n, n4) < 0.4) if dropout else None
    ref = R.forward(p, s, a, n3, n4, m3, m4)
    kw = dict(mask3=T_(m3, dev, torch.uint8), mask4=T_(m4, dev, torch.uint8)) if dropout else {}
    r = engine.rnet_forward(T_(p, dev), T_(s, dev), T_(a, dev), n3, n4, **kw).cpu().numpy()
    np.testing.assert_allclose(r, ref, rtol=1e-5, atol=1e-6)
    assert engine.rnet_param_count(d, n3, n4) == R.param_count(d, n3, n4)
    assert engine.rnet_param_offsets(d, n3, n4) == [off for _, _, off in R.layout(d, n3, n4)]


@pytest.mark.parametrize("d,n3,n4", [(15, 8, 4), (4, 6, 8), (16, 8, 8), (15, 6, 6)])
@pytest.mark.parametrize("dropout", [False, True])
def test_backward_matches_oracle(dev, d, n3, n4, dropout):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(d * 10 + n4)
    n = 777
    p, s, a = make(rng, n, d, n3, n4, scale=0.1)
    m3 = (rng.rand(n, n3) < 0.4) if dropout else None
    m4 = (rng.rand(n, n4) < 0.4) if dropout else None
    dr = f32(rng.randn(n))
    r_ref, cache = R.forward(p, s, a, n3, n4, m3, m4, cache=True)
    g_ref = R.backward(cache, dr)
    kw = dict(mask3=T_(m3, dev, torch.uint8), mask4=T_(m4, dev, torch.uint8)) if dropout else {}
    g, r = engine.rnet_backward(T_(p, dev), T_(s, dev), T_(a, dev), T_(dr, dev), n3, n4, want_rewards=True, **kw)
    np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-5, atol=1e-6)
    g = g.cpu().numpy()
    # per-tensor scale: a float32 sum over 777 transitions of mixed-sign terms
    for name, shp, off in R.layout(d, n3, n4):
        sl = slice(off, off + int(np.prod(shp)))
        scale = np.abs(g_ref[sl]).max() + 1e-6
        assert np.abs(g[sl] - g_ref[sl]).max() <= 1e-5 * scale + 1e-6, name
    # accumulate=True adds, and the reduction is deterministic
    g2 = engine.rnet_backward(T_(p, dev), T_(s, dev), T_(a, dev), T_(dr, dev), n3, n4,
                              grad=T_(g, dev).clone(), accumulate=True, **kw).cpu().numpy()
    np.testing.assert_array_equal(g2, f32(g) + f32(g))


def test_philox_dropout_is_reproducible_and_has_the_right_rate(dev):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(5)
    d, n3, n4, n = 15, 8, 4, 4096
    p, s, a = make(rng, n, d, n3, n4, scale=0.3)
    P, S, A = T_(p, dev), T_(s, dev), T_(a, dev)
    r1 = engine.rnet_forward(P, S, A, n3, n4, seed=11)
    r2 = engine.rnet_forward(P, S, A, n3, n4, seed=11)
    r3 = engine.rnet_forward(P, S, A, n3, n4, seed=12)
    assert torch.equal(r1, r2) and not torch.equal(r1, r3)
    # shifting the sample offset shifts the masks with the data
    r4 = engine.rnet_forward(P, S[100:], A[100:], n3, n4, seed=11, sample_offset=100)
    assert torch.equal(r1[100:], r4)
    # with a zero output bias, all fc4 units dropped <=> r == 0: rate (1-keep)^n4 = 0.6^4
    off = engine.rnet_param_offsets(d, n3, n4)
    p0 = p.copy()
    p0[off[9]] = 0.0
    p0[off[7]:off[7] + n4] = 1.0                        # fc4 biases > 0 so relu never kills the unit
    p0[off[6]:off[7]] = np.abs(p0[off[6]:off[7]])
    p0[off[8]:off[8] + n4] = 1.0
    r0 = engine.rnet_forward(T_(p0, dev), S, A, n3, n4, seed=3).cpu().numpy()
    frac = float((r0 == 0).mean())
    assert abs(frac - 0.6 ** 4) < 4 * np.sqrt(0.13 * 0.87 / n)


@pytest.mark.parametrize("layout", ["time_major", "trajectory_major"])
@pytest.mark.parametrize("with_z", [False, True])
def test_irl_loss_and_derivatives(dev, layout, with_z):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(6)
    M, T, nd = 301, 15, 75
    rd = f32(rng.uniform(-1, 1, nd))
    rg = f32(rng.uniform(-1, 1, (M, T)))                # trajectory-major reference view
    lz = f32(rng.randn(M)) if with_z else None
    first, second, dd, dg = R.irl_loss(rd, rg, 5, log_z=lz)
    dev_rg = T_(rg if layout == "trajectory_major" else rg.T, dev)
    res = engine.irl_loss_grad(T_(rd, dev), dev_rg, T, 5, layout=layout, log_z=None if lz is None else T_(lz, dev))
    loss = res["loss"].cpu().numpy()
    np.testing.assert_allclose(loss[:3], [first + second, first, second], rtol=1e-12)
    np.testing.assert_allclose(res["d_demo"].cpu().numpy(), dd, rtol=1e-7)
    got = res["d_gen"].cpu().numpy()
    got = got if layout == "trajectory_major" else got.T
    np.testing.assert_allclose(got, dg, rtol=1e-6)


@pytest.mark.parametrize("layout", ["time_major", "trajectory_major"])
@pytest.mark.parametrize("d,n3,n4,M,T,dropout", [(15, 8, 4, 53, 15, False), (15, 8, 4, 300, 15, True), (15, 6, 6, 7, 16, False),
                                                 (4, 6, 8, 41, 3, True), (16, 8, 8, 1, 1, False)])
def test_one_pass_generated_half_matches_oracle(dev, layout, d, n3, n4, M, T, dropout):
    """dmfg_rnet_backward_gen (forward, R_j, backward with weight exp(R_j), 1/Z on the reduced gradient) against the
    float64 oracle of the three-step chain it replaces (forward -> ac_irl.py:390-406 loss and dL/dr -> backward):
    loss terms, r_gen and the gradient accumulated on top of an existing buffer; ragged trajectory counts, T up to
    the 16 groups of a CTA, injected dropout masks, both layouts."""
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(d * 100 + M)
    n = M * T
    p, s, a = make(rng, n, d, n3, n4, scale=0.3)                 # [M, T] trajectory-major reference view
    m3 = (rng.rand(n, n3) < 0.4) if dropout else None
    m4 = (rng.rand(n, n4) < 0.4) if dropout else None
    rd = f32(rng.uniform(-1, 1, 75))
    r_ref, cache = R.forward(p, s, a, n3, n4, m3, m4, cache=True)
    first, second, _, dg = R.irl_loss(rd, r_ref.reshape(M, T), 5)
    g_ref = R.backward(cache, dg.reshape(-1))

    def order(x):                                                # trajectory-major [M*T, ...] -> the device layout
        if x is None or layout == "trajectory_major":
            return x
        return np.ascontiguousarray(x.reshape((M, T) + x.shape[1:]).swapaxes(0, 1)).reshape(x.shape)

    kw = dict(mask3=T_(order(m3), dev, torch.uint8), mask4=T_(order(m4), dev, torch.uint8)) if dropout else {}
    g0 = f32(rng.randn(p.size))
    g, loss, r = engine.rnet_backward_gen(T_(p, dev), T_(order(s), dev), T_(order(a), dev), n3, n4, T, T_(rd, dev), 5,
                                          layout=layout, grad=T_(g0, dev), accumulate=True, want_rewards=True, **kw)
    np.testing.assert_allclose(r.cpu().numpy(), order(r_ref), rtol=1e-5, atol=1e-6)
    # R_j is a float sum of T float rewards: |err| ~ T * 2e-6 enters ln Z
    np.testing.assert_allclose(loss.cpu().numpy()[:3], [first + second, first, second], rtol=1e-5, atol=3e-5)
    g = g.cpu().numpy() - g0
    for name, shp, off in R.layout(d, n3, n4):
        sl = slice(off, off + int(np.prod(shp)))
        scale = np.abs(g_ref[sl]).max() + 1e-6
        assert np.abs(g[sl] - g_ref[sl]).max() <= 1e-5 * scale + 1e-6, name


def test_one_pass_generated_half_argument_errors(dev):
    from discrete_mean_field_game_b200 import engine
    from discrete_mean_field_game_b200._lib import DmfgError
    rng = np.random.RandomState(0)
    p, s, a = make(rng, 34, 15, 8, 4)
    rd = T_(f32(rng.rand(5)), dev)
    with pytest.raises(DmfgError, match="T <= 16"):
        engine.rnet_backward_gen(T_(p, dev), T_(s, dev), T_(a, dev), 8, 4, 17, rd, 5)
    with pytest.raises(ValueError):
        engine.rnet_backward_gen(T_(p, dev), T_(s, dev), T_(a, dev), 8, 4, 15, rd, 5)


@pytest.mark.parametrize("l1l2", [False, True])
def test_adam_tf_steps(dev, l1l2):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(7)
    d, n3, n4 = 15, 8, 4
    n = R.param_count(d, n3, n4)
    p = f32(rng.randn(n) * 0.1)
    m = np.zeros(n)
    v = np.zeros(n)
    P, Mm, V = T_(p, dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pr = p.astype(np.float64)
    for step in range(1, 6):
        g = f32(rng.randn(n))
        gr = g.astype(np.float64) + (R.reg_grad(pr, d, n3, n4) if l1l2 else 0)
        reg_ref = R.reg_loss(pr, d, n3, n4)
        pr, m, v = R.adam_tf(pr, m, v, gr, step, 1e-4)
        reg = engine.adam_tf(P, Mm, V, T_(g, dev), step, 1e-4, l1l2=l1l2, net=(d, n3, n4), want_reg_loss=True)
        np.testing.assert_allclose(float(reg[0]), reg_ref, rtol=1e-6)
        np.testing.assert_allclose(P.cpu().numpy(), pr, rtol=0, atol=2e-7)
    np.testing.assert_allclose(Mm.cpu().numpy(), m, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(V.cpu().numpy(), v, rtol=1e-5, atol=1e-9)


def test_dirichlet_logq_and_log_z(dev):
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(8)
    d, M, T, K = 15, 7, 15, 10
    s = f32(rng.dirichlet(np.ones(d), size=(M, T)))
    a = f32(rng.dirichlet(np.ones(d) * 2, size=(M, T, d)))
    thetas = rng.uniform(6, 9, K)
    lq_ref = R.log_q(s, a, thetas, 0.0)
    lz_ref = R.log_z(s, a, thetas, 0.0, 21)
    # trajectory-major rows n = j*T + t
    lq = engine.dirichlet_logq(T_(s.reshape(M * T, d), dev), T_(a.reshape(M * T, d, d), dev),
                               T_(thetas, dev, torch.float64), 0.0)
    np.testing.assert_allclose(lq.cpu().numpy().reshape(M, T, K).sum(1), lq_ref, rtol=1e-9)
    lz = engine.irl_log_z(lq, T, 21, layout="trajectory_major").cpu().numpy()
    np.testing.assert_allclose(lz, lz_ref, rtol=1e-6)
    # time-major rows n = t*M + j
    lq_t = engine.dirichlet_logq(T_(s.transpose(1, 0, 2).reshape(M * T, d), dev),
                                 T_(a.transpose(1, 0, 2, 3).reshape(M * T, d, d), dev),
                                 T_(thetas, dev, torch.float64), 0.0)
    lz_t = engine.irl_log_z(lq_t, T, 21, layout="time_major").cpu().numpy()
    np.testing.assert_allclose(lz_t, lz_ref, rtol=1e-6)


def test_rnet_argument_errors(dev):
    from discrete_mean_field_game_b200 import engine
    from discrete_mean_field_game_b200._lib import DmfgError
    rng = np.random.RandomState(9)
    p, s, a = make(rng, 4, 15, 8, 4)
    with pytest.raises(DmfgError) as e:
        engine.rnet_forward(torch.zeros(R.param_count(15, 12, 4), device=dev), T_(s, dev), T_(a, dev), 12, 4)
    assert e.value.code == -2                                   # UNSUPPORTED: n_fc3 > 8
    r = engine.rnet_forward(T_(p, dev), T_(s[:0], dev), T_(a[:0], dev), 8, 4)
    assert r.numel() == 0
    g = engine.rnet_backward(T_(p, dev), T_(s[:0], dev), T_(a[:0], dev), torch.zeros(0, device=dev), 8, 4)
    assert float(g.abs().sum()) == 0.0


def test_random_shapes_forward_backward_vs_oracle(dev):
    """12 random (d, n_fc3, n_fc4, N, dropout) shapes -- every d up to 16, widths 1..8, ragged N down to a single
    transition, sparse actions -- through the specialised (d = 15) and the generic instantiations of the kernels."""
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(31)
    for trial in range(12):
        d = int(rng.choice([15, 15, 16, 3, 8, 11]))
        n3, n4 = int(rng.randint(1, 9)), int(rng.randint(1, 9))
        n = int(rng.choice([1, 15, 17, 100, 333]))
        p = f32(R.xavier_init(d, n3, n4, rng) + 0.1 * rng.randn(R.param_count(d, n3, n4)))
        s = f32(rng.dirichlet(np.ones(d) * 0.3, size=n))
        a = f32(rng.dirichlet(np.ones(d) * float(rng.choice([0.05, 0.5])), size=(n, d)))
        dropout = bool(trial % 2)
        m3 = (rng.rand(n, n3) < 0.4) if dropout else None
        m4 = (rng.rand(n, n4) < 0.4) if dropout else None
        dr = f32(rng.randn(n))
        r_ref, cache = R.forward(p, s, a, n3, n4, m3, m4, cache=True)
        g_ref = R.backward(cache, dr)
        kw = dict(mask3=T_(m3, dev, torch.uint8), mask4=T_(m4, dev, torch.uint8)) if dropout else {}
        tag = "trial %d: d=%d n3=%d n4=%d N=%d dropout=%s" % (trial, d, n3, n4, n, dropout)
        r_f = engine.rnet_forward(T_(p, dev), T_(s, dev), T_(a, dev), n3, n4, **kw).cpu().numpy()
        g, r_b = engine.rnet_backward(T_(p, dev), T_(s, dev), T_(a, dev), T_(dr, dev), n3, n4, want_rewards=True, **kw)
        np.testing.assert_allclose(r_f, r_ref, rtol=1e-5, atol=1e-6, err_msg=tag)
        np.testing.assert_array_equal(r_b.cpu().numpy(), r_f, err_msg=tag)
        g = g.cpu().numpy()
        for name, shp, off in R.layout(d, n3, n4):
            sl = slice(off, off + int(np.prod(shp)))
            scale = np.abs(g_ref[sl]).max() + 1e-6
            assert np.abs(g[sl] - g_ref[sl]).max() <= 1e-5 * scale + 1e-6, tag + " " + name


@pytest.mark.parametrize("d", [20, 21])
def test_reward_net_d20_d21_forward_backward_vs_oracle(dev, d):
    """The 32-lanes-per-transition instantiations: d = 20 (the reference's action files are 20 x 20, ac_irl.py:164-200)
    and d = 21 (mfg_ac2.py:25's default), forward and backward (fc3 weight gradient on the tensor cores with 8 M-tiles)."""
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(d)
    for n3, n4, n, dropout in ((8, 4, 77, False), (5, 7, 8, True), (8, 4, 1, False)):
        p = f32(R.xavier_init(d, n3, n4, rng) + 0.1 * rng.randn(R.param_count(d, n3, n4)))
        s = f32(rng.dirichlet(np.ones(d) * 0.3, size=n))
        a = f32(rng.dirichlet(np.ones(d) * 0.5, size=(n, d)))
        m3 = (rng.rand(n, n3) < 0.4) if dropout else None
        m4 = (rng.rand(n, n4) < 0.4) if dropout else None
        dr = f32(rng.randn(n))
        r_ref, cache = R.forward(p, s, a, n3, n4, m3, m4, cache=True)
        g_ref = R.backward(cache, dr)
        kw = dict(mask3=T_(m3, dev, torch.uint8), mask4=T_(m4, dev, torch.uint8)) if dropout else {}
        tag = "d=%d n3=%d n4=%d N=%d dropout=%s" % (d, n3, n4, n, dropout)
        r_f = engine.rnet_forward(T_(p, dev), T_(s, dev), T_(a, dev), n3, n4, **kw).cpu().numpy()
        g, r_b = engine.rnet_backward(T_(p, dev), T_(s, dev), T_(a, dev), T_(dr, dev), n3, n4, want_rewards=True, **kw)
        np.testing.assert_allclose(r_f, r_ref, rtol=1e-5, atol=1e-6, err_msg=tag)
        np.testing.assert_array_equal(r_b.cpu().numpy(), r_f, err_msg=tag)
        g = g.cpu().numpy()
        for name, shp, off in R.layout(d, n3, n4):
            sl = slice(off, off + int(np.prod(shp)))
            scale = np.abs(g_ref[sl]).max() + 1e-6
            assert np.abs(g[sl] - g_ref[sl]).max() <= 1e-5 * scale + 1e-6, tag + " " + name


def test_reward_net_forward_generic_wide_d_and_limits(dev):
    """16 < d <= 32 runs the generic 32-lane forward kernel; the backward kernel is built for d <= 16 and d = 20 / 21 and
    says so; d > 32 is refused."""
    from discrete_mean_field_game_b200 import engine
    from discrete_mean_field_game_b200._lib import DmfgError, ERR_UNSUPPORTED
    rng = np.random.RandomState(2)
    for d in (17, 27, 32):
        n = 9
        p = f32(R.xavier_init(d, 6, 3, rng))
        s = f32(rng.dirichlet(np.ones(d), size=n))
        a = f32(rng.dirichlet(np.ones(d), size=(n, d)))
        r = engine.rnet_forward(T_(p, dev), T_(s, dev), T_(a, dev), 6, 3).cpu().numpy()
        np.testing.assert_allclose(r, R.forward(p, s, a, 6, 3), rtol=1e-5, atol=1e-6)
    with pytest.raises(DmfgError) as e:
        engine.rnet_backward(T_(p, dev), T_(s, dev), T_(a, dev), T_(f32(rng.randn(n)), dev), 6, 3)
    assert e.value.code == ERR_UNSUPPORTED
    d = 33
    with pytest.raises(DmfgError) as e:
        engine.rnet_forward(T_(f32(R.xavier_init(d, 6, 3, rng)), dev), T_(f32(rng.dirichlet(np.ones(d), size=2)), dev),
                            T_(f32(rng.dirichlet(np.ones(d), size=(2, d))), dev), 6, 3)
    assert e.value.code == ERR_UNSUPPORTED


@pytest.mark.parametrize("d,n3,n4", [(15, 8, 4), (4, 6, 8), (16, 8, 8)])
def test_gathered_batch_equals_the_stacked_batch(dev, d, n3, n4):
    """dmfg_rnet_args.gather_*: the batch as slot numbers into a pool of resident trajectories gives bit for bit what the
    same trajectories stacked into [N, d] / [N, d, d] arrays give -- forward, backward (gradient and rewards), with the
    Philox dropout offsets following the position in the batch, not in the pool."""
    from discrete_mean_field_game_b200 import engine
    rng = np.random.RandomState(3 * d + n3)
    Tt, slots_total = 15, 40
    p, s, a = make(rng, slots_total * Tt, d, n3, n4, scale=0.2)
    P, S, A = T_(p, dev), T_(s, dev), T_(a, dev)
    slots = [int(x) for x in rng.permutation(slots_total)[:11]] + [7, 7]        # repeats are allowed
    rows = np.concatenate([np.arange(sl * Tt, (sl + 1) * Tt) for sl in slots])
    Ss, As = S[torch.as_tensor(rows, device=dev)].contiguous(), A[torch.as_tensor(rows, device=dev)].contiguous()
    dr = T_(f32(rng.randn(len(rows))), dev)
    for kw in ({}, dict(seed=11, sample_offset=1000)):
        r0 = engine.rnet_forward(P, Ss, As, n3, n4, **kw)
        r1 = engine.rnet_forward(P, S, A, n3, n4, gather=(Tt, slots), **kw)
        assert torch.equal(r0, r1)
        g0, q0 = engine.rnet_backward(P, Ss, As, dr, n3, n4, want_rewards=True, **kw)
        g1, q1 = engine.rnet_backward(P, S, A, dr, n3, n4, want_rewards=True, gather=(Tt, slots), **kw)
        assert torch.equal(g0, g1) and torch.equal(q0, q1)
    with pytest.raises(ValueError):
        engine.rnet_forward(P, S, A, n3, n4, gather=(Tt, [slots_total]))          # outside the pool
    with pytest.raises(ValueError):
        engine.rnet_forward(P, S, A, n3, n4, gather=(Tt, list(range(33))))        # more than DMFG_MAX_GATHER
