"""Restatement of the Masksembles mask generator (test infrastructure, see oracle/__init__.py).

Follows ``Software_Artifact/software/utils.py``: ``generate_masks_`` :18-41,
``generate_masks`` :44-63, ``generation_wrapper`` :66-110.  The reference consumes the GLOBAL
NumPy RNG (``np.random.choice(range(total), m, replace=False)`` once per mask, :34), so
bit-exactness means: same sequence of global-RNG calls with the same arguments.  Pinned
against the live reference by ``tests/golden/masksembles_masks.npz``.
"""
import numpy as np


def _draw_once(ones, n, s):
    # utils.py:29-41 - n random rows with `ones` ones among int(ones*s) positions, then drop
    # the columns no row uses.
    width = int(ones * s)
    rows = np.zeros((n, width))
    for r in range(n):
        hit = np.random.choice(range(width), ones, replace=False)
        rows[r, hit] = 1
    used = rows.any(axis=0)
    return rows[:, used]


def _draw_until_expected(ones, n, s):
    # utils.py:58-63 - redraw until the surviving width equals the closed-form expectation.
    want = int(ones * s * (1 - (1 - 1 / s) ** n))
    out = _draw_once(ones, n, s)
    while out.shape[1] != want:
        out = _draw_once(ones, n, s)
    return out


def generation_wrapper(c, n, scale):
    """float64 [n, c] matrix of {0,1}; raises ValueError like utils.py:77-85,106-108."""
    if c < 10:
        raise ValueError("Masksembles needs at least 10 channels, got channels=%d" % c)
    if scale > 6.0:
        raise ValueError("Masksembles scale must be <= 6, got scale=%s" % scale)
    ones = int(int(c) / (scale * (1 - (1 - 1 / scale) ** n)))       # utils.py:88
    masks = _draw_until_expected(ones, n, scale)                     # utils.py:93
    s = None
    for s in np.linspace(max(0.8 * scale, 1.0), 1.5 * scale, 300):   # utils.py:94-98
        if masks.shape[-1] >= c:
            break
        masks = _draw_until_expected(ones, n, s)
    upper = s
    if masks.shape[-1] != c:                                         # utils.py:100-104
        for s in np.linspace(max(0.8 * scale, 1.0), upper, 1000):
            if masks.shape[-1] >= c:
                break
            masks = _draw_until_expected(ones, n, s)
    if masks.shape[-1] != c:
        raise ValueError("generation_wrapper failed to produce %d features; change scale" % c)
    return masks
