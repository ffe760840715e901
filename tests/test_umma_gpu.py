"""The tcgen05 / TMEM path of the reward net's fc3 weight gradient in isolation (dmfg_umma_selftest): operand tiles in
the un-swizzled canonical layouts (A = h^T MN-major with a 144-byte chunk pitch, B = dz3 MN-major), kind::tf32 MMAs with
the 3xTF32 split, accumulation in TMEM across passes, commit -> mbarrier hand-off, TMEM read-back."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

eng = pytest.importorskip("discrete_mean_field_game_b200.engine")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    eng.require_cuda()
    return torch.device("cuda:0")


@pytest.mark.parametrize("passes", [1, 2, 7])
def test_umma_w3_gradient_tiles_match_float64(dev, passes):
    rng = np.random.RandomState(passes)
    h = np.float32(np.maximum(rng.randn(passes, 16, 512), 0.0) * rng.rand(passes, 16, 512))     # relu-like activations
    z = np.float32(rng.randn(passes, 16, 8) * 10.0 ** rng.randint(-3, 2, size=(passes, 16, 1)))
    out = eng.umma_selftest(torch.as_tensor(h, device=dev), torch.as_tensor(z, device=dev)).double().cpu().numpy()
    ref = np.einsum("pnk,pnj->kj", h.astype(np.float64), z.astype(np.float64))
    mag = np.einsum("pnk,pnj->kj", np.abs(h).astype(np.float64), np.abs(z).astype(np.float64))
    # 3xTF32: the dropped lo*lo term is 2^-22 of each product; FP32 accumulation adds ~1e-7 per term
    assert np.all(np.abs(out - ref) <= 2e-6 * mag + 1e-30), float(np.max(np.abs(out - ref) / (mag + 1e-30)))
    # a plain TF32 product (no split) would be off by ~5e-4: make sure the split is really in effect
    assert np.max(np.abs(out - ref) / (mag + 1e-30)) < 1e-5


def test_umma_structured_inputs_locate_every_element(dev):
    """h[n, k] = 1 only at (n0, k0), z[n, j] = n + 10 j: out[k0, j] = z[n0, j] and zero elsewhere, for every tile corner."""
    for n0, k0 in ((0, 0), (7, 3), (8, 127), (15, 128), (3, 449), (9, 511), (5, 258)):
        h = np.zeros((1, 16, 512), np.float32)
        h[0, n0, k0] = 1.0
        z = np.float32(np.arange(16)[:, None] + 10.0 * np.arange(8)[None, :])[None]
        out = eng.umma_selftest(torch.as_tensor(h, device=dev), torch.as_tensor(z, device=dev)).cpu().numpy()
        ref = np.zeros((512, 8), np.float32)
        ref[k0] = z[0, n0]
        np.testing.assert_array_equal(out, ref, err_msg="(n0, k0) = (%d, %d)" % (n0, k0))
