"""Shared helpers for the test-suite (fixture loading, tolerances)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_params(g):
    """The non-default model parameters a *_params fixture was recorded with (dict for OracleVecEnv / QuadVecEnv), or None."""
    prm = {k[len("param_"):]: (v.tolist() if v.ndim else float(v)) for k, v in g.items() if k.startswith("param_")}
    return prm or None


def close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float(np.max(np.abs(a - b))) if a.size else 0.0
    assert err <= tol, "max abs err %.3e > %.1e" % (err, tol)
    return err
