"""Shared helpers for the parity tests."""

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    """Return (case, ref) dicts from tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    case, ref = {}, {}
    for k in z.files:
        v = z[k]
        if v.ndim == 0:
            v = v.item()
        (case if k.startswith('in_') else ref)[k.split('_', 1)[1]] = v
    return case, ref


def max_rel(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over finite entries; NaN patterns must agree."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    den = np.maximum(np.abs(b[m]), floor if floor > 0 else 1e-300)
    return float(np.max(np.abs(a[m] - b[m]) / den))


def bitwise_equal(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == 'f':
        return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    return np.array_equal(a, b)
