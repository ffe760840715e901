"""The reward-net / IRL-loss / Adam / calc_z oracle against independent implementations (CPU).

The TF boundary is unpinned (no TensorFlow here), so the restatement is checked
against torch.nn.functional + autograd (a second, independently written implementation of
networks.py:13-157 and ac_irl.py:382-418), finite differences, torch.optim-free Adam algebra and the
hand-rolled Dirichlet pdf of the reference's test_acirl.py:15-30.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as Fnn
from scipy import special

from oracle import rnet_oracle as R

D, N3, N4, T = 15, 8, 4, 15


def torch_rnet(p, states, actions, d, n3, n4, mask3=None, mask4=None, keep=0.4):
    """networks.py:23-43 written with torch ops (NCHW conv, then permute back to NHWC to flatten)."""
    lay = {name: (shp, off) for name, shp, off in R.layout(d, n3, n4)}

    def get(name):
        shp, off = lay[name]
        return p[off:off + int(np.prod(shp))].reshape(shp)
    x = actions.reshape(-1, 1, d, d)
    w1 = get("conv1/weights").permute(3, 2, 0, 1)
    c1 = torch.relu(Fnn.conv2d(x, w1, get("conv1/biases"), padding=2))
    w2 = get("conv2/weights").permute(3, 2, 0, 1)
    c2 = torch.relu(Fnn.conv2d(c1, w2, get("conv2/biases"), padding=1))
    flat = c2.permute(0, 2, 3, 1).reshape(-1, 2 * d * d)
    h3 = torch.relu(flat @ get("fc3/weights") + get("fc3/biases"))
    if mask3 is not None:
        h3 = h3 * mask3 / keep
    h4 = torch.relu(torch.cat([h3, states], 1) @ get("fc4/weights") + get("fc4/biases"))
    if mask4 is not None:
        h4 = h4 * mask4 / keep
    return torch.tanh(h4 @ get("out/weights") + get("out/biases"))[:, 0]


def make_batch(rng, n, d=D):
    states = rng.dirichlet(np.ones(d), size=n)
    actions = rng.dirichlet(np.ones(d) * 0.5, size=(n, d))
    return states, actions


def test_layout_and_param_count():
    assert R.param_count(15, 8, 4) == 3755
    lay = R.layout(15, 8, 4)
    assert [n for n, _, _ in lay][:4] == ["conv1/weights", "conv1/biases", "conv2/weights", "conv2/biases"]
    assert lay[4][1] == (450, 8) and lay[6][1] == (23, 4)
    p = R.xavier_init(15, 8, 4, np.random.RandomState(0))
    u = R.unpack(p, 15, 8, 4)
    assert np.all(u["fc3/biases"] == 0) and np.abs(u["fc3/weights"]).max() <= math.sqrt(6 / 458)


@pytest.mark.parametrize("dropout", [False, True])
@pytest.mark.parametrize("d,n3,n4", [(15, 8, 4), (4, 6, 8), (16, 4, 6)])
def test_forward_backward_match_torch_autograd(d, n3, n4, dropout):
    rng = np.random.RandomState(1)
    p = R.xavier_init(d, n3, n4, rng) + 0.05 * rng.randn(R.param_count(d, n3, n4))
    s, a = make_batch(rng, 37, d)
    m3 = (rng.rand(37, n3) < 0.4).astype(np.float64) if dropout else None
    m4 = (rng.rand(37, n4) < 0.4).astype(np.float64) if dropout else None
    dr = rng.randn(37)
    r, cache = R.forward(p, s, a, n3, n4, m3, m4, cache=True)
    g = R.backward(cache, dr)
    pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    rt = torch_rnet(pt, torch.tensor(s), torch.tensor(a), d, n3, n4,
                    None if m3 is None else torch.tensor(m3), None if m4 is None else torch.tensor(m4))
    np.testing.assert_allclose(r, rt.detach().numpy(), rtol=1e-12, atol=1e-14)
    (rt * torch.tensor(dr)).sum().backward()
    np.testing.assert_allclose(g, pt.grad.numpy(), rtol=1e-10, atol=1e-13)


def test_loss_and_grad_match_torch_and_finite_differences():
    rng = np.random.RandomState(2)
    p = R.xavier_init(D, N3, N4, rng) + 0.05 * rng.randn(3755)
    ds, da = make_batch(rng, 5 * T)
    gs, ga = make_batch(rng, 5 * T)
    for reg in ("none", "l1l2"):
        loss, (first, second), grad = R.loss_and_grad(p, ds, da, gs, ga, N3, N4, 5, T, reg=reg)
        pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        rd = torch_rnet(pt, torch.tensor(ds), torch.tensor(da), D, N3, N4)
        rg = torch_rnet(pt, torch.tensor(gs), torch.tensor(ga), D, N3, N4)
        lt = -rd.sum() / 5 + torch.log(torch.exp(rg.reshape(5, T).sum(1)).sum() / 5)     # ac_irl.py:390-406
        if reg == "l1l2":
            m = torch.tensor(R.reg_mask(D, N3, N4))
            lt = lt + (pt.abs() * m).sum() + 0.5 * (pt * pt * m).sum()
        lt.backward()
        np.testing.assert_allclose(loss, float(lt.detach()), rtol=1e-12)
        np.testing.assert_allclose(grad, pt.grad.numpy(), rtol=1e-9, atol=1e-12)
    # finite differences on a few coordinates of every layer
    loss0, _, grad = R.loss_and_grad(p, ds, da, gs, ga, N3, N4, 5, T)
    for name, shp, off in R.layout(D, N3, N4):
        i = off + int(np.prod(shp)) // 2
        h = 1e-6
        pp, pm = p.copy(), p.copy()
        pp[i] += h
        pm[i] -= h
        fd = (R.loss_and_grad(pp, ds, da, gs, ga, N3, N4, 5, T)[0] - R.loss_and_grad(pm, ds, da, gs, ga, N3, N4, 5, T)[0]) / (2 * h)
        assert abs(fd - grad[i]) <= 1e-6 * max(1.0, abs(fd)), name


def test_irl_loss_terms_and_weights():
    rng = np.random.RandomState(3)
    rd, rg = rng.uniform(-1, 1, 75), rng.uniform(-1, 1, (5, 15))
    first, second, dd, dg = R.irl_loss(rd, rg, 5)
    assert np.isclose(first, -rd.sum() / 5)                                   # quirk C.9
    assert np.isclose(second, math.log(np.mean(np.exp(rg.sum(1)))))
    assert np.allclose(dd, -0.2) and np.allclose(dg.sum(0), 1.0)
    lz = rng.randn(5)
    _, second_z, _, dgz = R.irl_loss(rd, rg, 5, log_z=lz)
    assert np.isclose(second_z, math.log(np.mean(np.exp(lz) * np.exp(rg.sum(1)))))     # ac_irl.py:405
    w = np.exp(lz + rg.sum(1))
    assert np.allclose(dgz[:, 0], w / w.sum())


def test_one_pass_form_of_the_generated_gradient():
    """The algebra dmfg_rnet_backward_gen rests on: d second / d params = sum_j softmax_j(R) dR_j/dparams equals the
    gradient backpropagated with the UNNORMALISED weights exp(R_j), divided by Z = sum_j exp(R_j) afterwards, and
    second = ln(Z / M) (ac_irl.py:396-406 with z_j = 1).  |r| < 1 bounds exp(R_j) by e^T: no shift is needed."""
    rng = np.random.RandomState(8)
    d, n3, n4, M, Tt = 6, 5, 3, 7, 4
    p = R.xavier_init(d, n3, n4, rng) + 0.3 * rng.randn(R.param_count(d, n3, n4))
    s = rng.dirichlet(np.ones(d), size=M * Tt)
    a = rng.dirichlet(np.ones(d) * 0.5, size=(M * Tt, d))
    r, cache = R.forward(p, s, a, n3, n4, None, None, cache=True)
    assert np.all(np.abs(r) < 1)
    _, second, _, dg = R.irl_loss(np.zeros(1), r.reshape(M, Tt), 1)
    g_ref = R.backward(cache, dg.reshape(-1))
    u = np.exp(r.reshape(M, Tt).sum(1))
    g_u = R.backward(cache, np.repeat(u, Tt))
    np.testing.assert_allclose(g_u / u.sum(), g_ref, rtol=1e-11, atol=1e-15)
    assert np.isclose(second, math.log(u.sum() / M), rtol=1e-13)


def test_adam_tf_first_steps():
    rng = np.random.RandomState(4)
    p, g = rng.randn(10), rng.randn(10)
    m = v = np.zeros(10)
    p1, m1, v1 = R.adam_tf(p, m, v, g, 1, 1e-4)
    # t = 1: lr_t = lr*sqrt(1-b2)/(1-b1); m = (1-b1) g; v = (1-b2) g^2
    lr_t = 1e-4 * math.sqrt(0.001) / 0.1
    np.testing.assert_allclose(p1, p - lr_t * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-8), rtol=1e-12)
    # differs from torch.optim.Adam (eps inside the bias correction) only through eps
    p2, _, _ = R.adam_tf(p1, m1, v1, g, 2, 1e-4)
    assert np.all(np.abs(p2 - p1) < 1.01e-4) and np.all(np.sign(p1 - p2) == np.sign(g))


def ref_dirichlet(alpha, x):
    """test_acirl.py:15-30 (the reference's own hand-rolled Dirichlet pdf)."""
    num = math.gamma(sum(alpha))
    den = 1.0
    for a in alpha:
        den *= math.gamma(a)
    val = num / den
    for a, xi in zip(alpha, x):
        val *= xi ** (a - 1)
    return val


def test_log_q_against_reference_hand_rolled_pdf():
    rng = np.random.RandomState(5)
    d, M, K, Tt = 4, 3, 2, 2
    s = rng.dirichlet(np.ones(d), size=(M, Tt))
    a = rng.dirichlet(np.ones(d) * 2, size=(M, Tt, d))
    thetas, shift = np.array([6.5, 8.0]), 0.1
    lq = R.log_q(s, a, thetas, shift)
    for j in range(M):
        for k in range(K):
            q = 1.0
            for t in range(Tt):
                diff = s[j, t][None, :] - s[j, t][:, None]
                alpha = np.maximum(np.log(1 + np.exp(thetas[k] * (diff - shift))), 1 + 1e-6)
                for i in range(d):
                    q *= ref_dirichlet(alpha[i], a[j, t, i])
            assert np.isclose(lq[j, k], math.log(q), rtol=1e-10)
    lz = R.log_z(s, a, thetas, shift, num_start_samples=21)
    assert np.allclose(np.exp(lz), K / (21 * np.exp(lq).sum(1)))                # ac_irl.py:379
