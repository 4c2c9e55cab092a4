"""Shared helpers of the test-suite."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def rel_err(a, b):
    """max|a - b| / max|b| (the parity metric of SURVEY.md section 8d)."""
    scale = np.max(np.abs(b))
    if scale == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)


def per_step_rel_err(a, b):
    return max(rel_err(a[i], b[i]) for i in range(len(b)))
