"""Shared comparison helpers for the parity tests."""

import numpy as np

# Parity gates: integer structures bit-exact; floating
# point tables within 1e-9 relative (absolute floor 1e-300).
RTOL = 1e-9
ATOL = 1e-300


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(all="ignore"):
        d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), ATOL)
    d[same] = 0.0
    d[np.isnan(d)] = np.inf
    return float(d.max())


def assert_close(a, b, name, rtol=RTOL):
    r = relerr(a, b)
    assert r <= rtol, "%s: max relative error %.3e > %.1e" % (name, r, rtol)


def first_divergence(p1, p2):
    p1 = np.asarray(p1)
    p2 = np.asarray(p2)
    bad = np.nonzero(p1 != p2)[0]
    return None if bad.size == 0 else int(bad.max())   # traceback runs backwards


def block_starts(blocklens):
    return np.concatenate([[0], np.cumsum(blocklens)]).astype(np.int64)
