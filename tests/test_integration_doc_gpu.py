"""INTEGRATION.md section 2 -- the torch-free ctypes binding a maintainer of the reference would add -- executed
verbatim: the C ABI works with plain host pointers (dmfg_rollout_host) and the struct in the document is current."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ctypes_stub_of_the_integration_guide_runs(monkeypatch):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    monkeypatch.chdir(ROOT)                                   # the stub opens the library by its in-tree path
    src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = [b for b in re.findall(r"```python\n(.*?)```", src, re.S) if "import ctypes" in b][0]
    ns = {}
    exec(code, ns)
    pi0 = np.random.RandomState(0).dirichlet(np.ones(15), size=5)
    S, A = ns["generate_trajectories"](pi0, 8.64, 0.0, 1e4)
    assert S.shape == (16, 5, 15) and A.shape == (15, 5, 15, 15)
    np.testing.assert_allclose(A.sum(-1), 1.0, atol=1e-6)
    np.testing.assert_allclose(np.einsum("tbi,tbij->tbj", S[:-1], A), S[1:], atol=3e-7)   # test_acirl.py:43-47
