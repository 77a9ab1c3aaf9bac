"""The window-parallel formulation of yakmo's sequential float prefix sum (tools/scan_model.py)
reproduces the sequential chain bit for bit -- including ties, tiny negative terms and binade crossings."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import scan_model as sm  # noqa: E402


def _weights(rng, n, kind):
    if kind == "distances":          # squared distances of a k-means++ step: wide range, some exact zeros
        a = (rng.standard_normal(n) * 0.05) ** 2 * np.exp(rng.standard_normal(n))
        a[rng.integers(0, n, n // 50)] = 0.0
        neg = rng.integers(0, n, n // 40)   # rounding of |p|^2 + |c|^2 - 2 p.c leaves tiny negative values
        a[neg] = -np.abs(rng.standard_normal(len(neg))) * 1e-9
    elif kind == "ties":             # multiples of a power of two: exact halves of the ulp all the time
        a = rng.integers(0, 64, n) * 2.0 ** -12
    else:                            # a few huge terms force crossings in the middle of windows
        a = np.abs(rng.standard_normal(n)) * 1e-3
        a[rng.integers(0, n, 6)] = 50.0
    return a.astype(np.float32)


def test_window_prefix_equals_sequential_chain():
    rng = np.random.default_rng(7)
    for kind in ("distances", "ties", "spikes"):
        a = _weights(rng, 6000, kind)
        ref = sm.seq_prefix(a)
        r0, exps, used0 = sm.window_prefix(a, None, win=512)          # no prediction: every window exact
        assert not any(used0) and np.array_equal(r0.view(np.uint32), ref.view(np.uint32))
        r1, _, used1 = sm.window_prefix(a, exps, win=512)             # perfect prediction
        assert np.array_equal(r1.view(np.uint32), ref.view(np.uint32)), kind
        assert sum(used1) >= len(used1) // 2, (kind, used1)           # most windows take the summary path


def test_prediction_from_previous_step_survives_decreasing_weights():
    """A seeding step only lowers some weights: the previous step's exponents stay a good (and always safe) guess."""
    rng = np.random.default_rng(11)
    a = _weights(rng, 8000, "distances")
    _, exps, _ = sm.window_prefix(a, None, win=512)
    total_used = 0
    for _ in range(4):
        idx = rng.integers(0, len(a), 200)
        a[idx] = (a[idx] * rng.random(len(idx))).astype(np.float32)   # some points got closer to the new seed
        ref = sm.seq_prefix(a)
        r, exps, used = sm.window_prefix(a, exps, win=512)
        assert np.array_equal(r.view(np.uint32), ref.view(np.uint32))
        total_used += sum(used)
    assert total_used > 0
