"""Runs the UNMODIFIED reference (``/root/reference/mfg_ac2.py``) when the checkout is present.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference is pure Python (NumPy + SciPy):
it is imported from where it lies, never copied.  On the GPU box the checkout does not exist, so
``available()`` is False there and ``bench.py --impl reference`` times the oracle port instead.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import tempfile
import time
import warnings

import numpy as np

REF = os.environ.get("DMFG_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "mfg_ac2.py"))


def import_reference(name="mfg_ac2"):
    """The reference promotes warnings to errors at import (mfg_ac2.py:21); reset the filters around it."""
    saved = warnings.filters[:]
    warnings.resetwarnings()
    warnings.simplefilter("ignore")
    sys.path.insert(0, REF)
    try:
        return importlib.import_module(name)
    finally:
        sys.path.remove(REF)
        warnings.filters[:] = saved


def write_start_states(dirname, rows):
    """trend_distribution_day<k>.csv, first line = start row, space separated %.3e (mfg_ac2.py:191-198)."""
    os.makedirs(dirname, exist_ok=True)
    for k, row in enumerate(rows, start=1):
        with open(os.path.join(dirname, "trend_distribution_day%d.csv" % k), "w") as f:
            f.write(" ".join("%.3e" % v for v in row) + "\n")


def time_train(episodes, d=15, theta=8.86349, shift=0.16, alpha_scale=12000.0, lr_critic=0.1, lr_actor=0.1, seed=0):
    """Seconds taken by ``mfg_ac2.actor_critic(...).train(num_episodes=episodes)`` of the reference itself
    (15 transitions per episode, mfg_ac2.py:478) on synthetic start states; returns (seconds, population_steps)."""
    from oracle.mfg_oracle import synthetic_start_states
    mod = import_reference("mfg_ac2")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        rows = synthetic_start_states(n_rows=21, n_cols=20, d=20, seed=0)
        write_start_states(os.path.join(tmp, "train_normalized_round2"), rows)
        os.chdir(tmp)
        try:
            np.random.seed(seed)
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ac = mod.actor_critic(theta=theta, shift=shift, alpha_scale=alpha_scale, d=d)
                t0 = time.perf_counter()
                ac.train(num_episodes=episodes, gamma=1, lr_critic=lr_critic, lr_actor=lr_actor,
                         consecutive=10 ** 9, write_file=0, write_all=0)
                dt = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    return dt, episodes * 15
