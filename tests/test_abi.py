"""CPU-side checks of the C ABI: the library loads, exports every symbol
include/dmfg.h declares, and its host-side pieces answer without a GPU."""
import ctypes as C
import os
import re

import pytest

from discrete_mean_field_game_b200 import _lib, build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dmfg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmfg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 12
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), "libdmfg.so does not export %s" % n
        assert n in bound, "%s is declared in dmfg.h but not bound in _lib.SYMBOLS" % n
    assert bound <= set(names)


def test_version_and_sizes(lib):
    assert lib.dmfg_version() == 1
    for d in (1, 3, 4, 15, 21, 47, 256):
        assert lib.dmfg_num_features(d) == d * (d + 1) // 2 + d + 1
        assert lib.dmfg_acc_len(d) == lib.dmfg_num_features(d) + 2
    assert lib.dmfg_num_features(15) == 136        # F = 136 at d = 15


def test_struct_size_guard(lib):
    a = _lib.RolloutArgs()
    a.struct_size = 8
    assert lib.dmfg_rollout(C.byref(a), None) == _lib.ERR_INVALID
    assert b"struct_size" in lib.dmfg_last_error()
    assert lib.dmfg_rollout(None, None) == _lib.ERR_INVALID


def test_argument_validation_without_gpu(lib):
    a = _lib.RolloutArgs()
    a.struct_size = C.sizeof(_lib.RolloutArgs)
    a.dtype, a.d, a.T, a.B = _lib.F32, 300, 1, 1
    assert lib.dmfg_rollout(C.byref(a), None) == _lib.ERR_INVALID
    a.d, a.alpha_scale, a.variant = 7, 1.0, _lib.VARIANT_FAST
    assert lib.dmfg_rollout(C.byref(a), None) == _lib.ERR_UNSUPPORTED
    a.variant = _lib.VARIANT_AUTO
    assert lib.dmfg_rollout(C.byref(a), None) == _lib.ERR_INVALID      # pi0 NULL
    assert b"pi0" in lib.dmfg_last_error()


def test_philox_known_answers(lib):
    """Random123 kat_vectors for philox4x32-10."""
    assert engine.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert engine.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert engine.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def _philox_numpy(ctr, key, rounds):
    """Independent restatement of Philox4x32-R (Salmon et al., SC'11) in Python integers."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(rounds):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & MASK, (p0 >> 32) ^ c3 ^ k1, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def test_gamma_stream_philox_rounds(lib):
    """The Gamma sampler's block function is Philox4x32-7: checked against an independent restatement that
    itself reproduces the published Random123 vectors at 10 rounds."""
    cases = [((0, 0, 0, 0), (0, 0)), ((0xffffffff,) * 4, (0xffffffff,) * 2),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)),
             ((12345, 0, 678, 3), (1234, 0))]
    assert _philox_numpy(*cases[0], 10) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _philox_numpy(*cases[2], 10) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    rounds = engine.gamma_philox_rounds()
    assert rounds == 7
    for ctr, key in cases:
        assert engine.philox(ctr, key) == _philox_numpy(ctr, key, 10)
        assert engine.philox(ctr, key, gamma_stream=True) == _philox_numpy(ctr, key, rounds)


def test_workspace_query(lib):
    a = _lib.RolloutArgs()
    a.struct_size = C.sizeof(_lib.RolloutArgs)
    a.dtype, a.d, a.T, a.B = _lib.F32, 15, 16, 1 << 20
    assert lib.dmfg_rollout_workspace_bytes(C.byref(a)) == 0            # nothing reduced, nothing needed
    a.w, a.acc = 1, 1                                                   # non-NULL markers (never dereferenced)
    fast = lib.dmfg_rollout_workspace_bytes(C.byref(a))
    assert 0 < fast < 4 << 20
    a.variant = _lib.VARIANT_GENERIC
    assert lib.dmfg_rollout_workspace_bytes(C.byref(a)) > fast          # generic records intermediates


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.require_cuda()
