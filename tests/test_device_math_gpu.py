"""Device math against SciPy: digamma accuracy and the statistical validation of the in-kernel
Gamma sampler (NumPy's MT19937 stream is not a parity target -- SURVEY 7 step 4)."""
import numpy as np
import pytest
import torch
from scipy import special, stats

pytestmark = pytest.mark.gpu

eng = pytest.importorskip("discrete_mean_field_game_b200.engine")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    eng.require_cuda()
    return torch.device("cuda:0")


def test_digamma_float64(dev):
    x = np.concatenate([np.logspace(-6, 3, 4000), np.linspace(0.5, 12, 2000), [1.8149927917809779e-02, 1.84e-5]])
    got = eng.digamma(torch.as_tensor(x, device=dev)).cpu().numpy()
    ref = special.digamma(x)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2)) < 1e-12


def test_digamma_float32(dev):
    """alpha spans ~[2e-5, 8] (App. B) and row sums reach ~1e2; psi is used in a sum, so the error that
    matters is absolute near the root at 1.46 and relative elsewhere."""
    x = np.concatenate([np.logspace(-6, 4.5, 6000), np.linspace(0.5, 12, 3000)]).astype(np.float32)
    got = eng.digamma(torch.as_tensor(x, device=dev)).cpu().numpy().astype(np.float64)
    ref = special.digamma(x.astype(np.float64))
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-6, err.max()


@pytest.mark.parametrize("shape", [0.05, 0.36, 0.999, 1.0, 2.5, 12.0, 217.0, 2600.0, 9.0e4])
def test_gamma_sampler_distribution(dev, shape):
    """KS test + first two moments of the Marsaglia-Tsang sampler (incl. the shape<1 boost)."""
    n = 400_000
    y = eng.gamma_sample(torch.full((n,), shape, device=dev), seed=11, pop=3).cpu().numpy().astype(np.float64)
    assert np.all(np.isfinite(y)) and np.all(y >= 0)
    dist = stats.gamma(a=shape)
    # float32 cannot hold the extreme lower tail of tiny shapes (Gamma(0.05) has 0.6 % of its mass
    # below 1e-45): check that censored mass, then KS on the representable part
    lo = 1e-30
    p_lo = dist.cdf(lo)
    n_lo = int((y < lo).sum())
    assert abs(n_lo - n * p_lo) < 6 * np.sqrt(n * p_lo * (1 - p_lo)) + 1
    ks = stats.kstest(y[y >= lo], lambda v: (dist.cdf(v) - p_lo) / (1 - p_lo))
    assert ks.pvalue > 1e-4, (shape, ks)
    se_mean = np.sqrt(shape / n)
    assert abs(y.mean() - shape) < 5 * se_mean
    var_se = np.sqrt((2 * shape ** 2 + 6 * shape) / n)        # var of sample variance of a gamma
    assert abs(y.var() - shape) < 6 * var_se


def test_gamma_sampler_mixed_shapes_and_streams(dev):
    """Neighbouring elements with very different shapes share one Philox call; distinct seeds /
    populations give distinct, uncorrelated streams; the same key reproduces bit for bit."""
    n = 200_000
    shapes = torch.tensor([0.3, 5000.0], device=dev).repeat(n // 2)
    a = eng.gamma_sample(shapes, seed=1, pop=0)
    b = eng.gamma_sample(shapes, seed=1, pop=0)
    c = eng.gamma_sample(shapes, seed=1, pop=1)
    e = eng.gamma_sample(shapes, seed=2, pop=0)
    assert torch.equal(a, b)
    an, cn, en = (t.cpu().numpy().astype(np.float64) for t in (a, c, e))
    for sh, sl in ((0.3, slice(0, None, 2)), (5000.0, slice(1, None, 2))):
        assert stats.kstest(an[sl], stats.gamma(a=sh).cdf).pvalue > 1e-4
        assert abs(np.corrcoef(an[sl], cn[sl])[0, 1]) < 0.02
        assert abs(np.corrcoef(an[sl], en[sl])[0, 1]) < 0.02
    # the two elements of a pair come from the two Box-Muller outputs of one call: uncorrelated
    assert abs(np.corrcoef(an[0::2], an[1::2])[0, 1]) < 0.02


def test_gamma_zero_shape(dev):
    y = eng.gamma_sample(torch.zeros(64, device=dev), seed=5)
    assert torch.count_nonzero(y) == 0          # np.random.gamma(0) == 0; the rollout maps it to 1e-20


@pytest.mark.parametrize("d", [15, 16, 21, 64])
def test_dirichlet_rows_from_rollout(dev, d):
    """P rows ~ Dirichlet(alpha * alpha_scale): E[P_ij] = alpha_ij / sum_j alpha_ij (mfg_ac2.py:238-249), through
    the v2 kernel (d = 15 / 16) and the wide kernel (odd d = 21, d = 64); alpha_scale = 50 puts many shapes below 1
    (boost path) and makes squeeze misses frequent (redo path)."""
    B = 1 << 15
    pi = np.random.RandomState(1).dirichlet(np.ones(d))
    pi0 = torch.as_tensor(np.repeat(pi[None], B, 0), dtype=torch.float32, device=dev)
    out = eng.rollout(pi0, 8.86349, 0.16, 50.0, 1, seed=3, outputs=("actions", "alpha", "alpha_deriv"))
    alpha = out["alpha"][0, 0].double().cpu().numpy()
    P = out["actions"][0].double().cpu().numpy()
    a = alpha * 50.0
    a0 = a.sum(-1, keepdims=True)
    mean, var = a / a0, a * (a0 - a) / (a0 ** 2 * (a0 + 1))
    assert np.all(np.abs(P.mean(0) - mean) < 6 * np.sqrt(var / B) + 1e-7)
    assert np.all(np.abs(P.var(0) - var) < 0.1 * var + 1e-9)
