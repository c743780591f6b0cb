"""Golden-fixture access for the tests (tests/golden/*.npz, written by tests/golden/make_fixtures.py from
the real reference) and the parity metrics of BASELINE.json's north_star:
  * primary-ray hit ids agree on >= 99.99 % of pixels (grazing-edge ties excepted),
  * 8-bit channels differ by at most 1 LSB on >= 99.9 % of pixels.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
VARIANT_KEYS = ('params', 'materials', 'point_lights', 'global_lights', 'ambient', 'bg1', 'bg2', 'bg3')


def load(name):
    """-> (scene dict, golden dict without the g_ prefix)"""
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        sc = {k: z[k] for k in z.files if not k.startswith('g_')}
        g = {k[2:]: z[k] for k in z.files if k.startswith('g_')}
    return sc, g


def variant(sc, g, name):
    """Scene dict with the non-geometry arrays of variant `name` swapped in."""
    out = dict(sc)
    for k in VARIANT_KEYS:
        out[k] = g['v_%s_%s' % (name, k)]
    if 'v_%s_simplex_mat' % name in g:
        out['simplex_mat'] = g['v_%s_simplex_mat' % name]
    return out


def quant8(rgb):
    """What an 8-bit channel of the reference's pixel packer stores (render.cpp:427,439)."""
    return np.floor(np.clip(rgb.astype(np.float64), 0, 1) * 255 + 0.5).astype(np.int32)


def lsb_stats(a, b, exclude=None):
    """fraction of pixels whose 8-bit channels differ by more than 1 LSB, and by more than 0."""
    d = np.abs(quant8(a) - quant8(b)).max(axis=-1)
    if exclude is not None:
        d = d[~exclude]
    n = max(d.size, 1)
    return float(np.count_nonzero(d > 1)) / n, float(np.count_nonzero(d > 0)) / n


def id_agreement(a, b, dist_a=None, dist_b=None, tie_tol=2e-5):
    """fraction of pixels with equal hit ids; pixels where both hit at (numerically) the same distance but
    report different primitives are exact-t ties between coincident facets and count as equal."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    same = a == b
    if dist_a is not None and dist_b is not None:
        da, db = np.asarray(dist_a).ravel(), np.asarray(dist_b).ravel()
        tie = (~same) & (a >= 0) & (b >= 0) & (np.abs(da - db) <= tie_tol * np.maximum(1.0, np.abs(da)))
        return float(np.count_nonzero(same | tie)) / a.size, int(np.count_nonzero(tie))
    return float(np.count_nonzero(same)) / a.size, 0


def center_column_mask(w, h):
    """x == w/2 is where polytope scenes have systematic exact-t ties."""
    m = np.zeros((h, w), dtype=bool)
    m[:, w // 2] = True
    return m
