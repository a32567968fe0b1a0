"""Shared helpers for the tests: golden-fixture access and tolerance checks."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def dense_grid(idx, val, G=84):
    g = np.zeros((G, G, 2), dtype=np.float64)
    if len(idx):
        g[idx[:, 0], idx[:, 1], idx[:, 2]] = val
    return g


def draws_of(d):
    return d["x0"], d["xa0"], d["burn"], d["agent_noise"], d["particle_noise"]


def rel_err(a, b):
    """max |a-b| / max|b|  -- the parity metric of SURVEY.md 7.3 (elementwise rtol with atol=0 is
    meaningless for just-landed locusts with y ~ 1e-4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
