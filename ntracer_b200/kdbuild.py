"""Host-side k-d tree builder behind build_kdtree / build_composite_scene (first cut).

Surface-area-heuristic splits over the prototypes' axis-aligned bounding boxes, sweep over candidate planes at
box boundaries, primitives overlapping the plane go to both sides (like the reference's tree, which this is NOT a
port of: src/tracer.hpp:1930-2455 clips with exact separating-axis tests and groups simplexes into SIMD batches).
Any correct tree gives the same hit ids and colours (shadows excepted)."""
import numpy as np

from . import tracern as T


def build(protos, max_depth=None, split_threshold=None, traversal_cost=None, intersection_cost=None):
    d = protos[0].dimension
    lo = np.stack([p.boundary.start._v for p in protos]).astype(np.float64)
    hi = np.stack([p.boundary.end._v for p in protos]).astype(np.float64)
    prims = [p.primitive for p in protos]
    n = len(prims)
    max_depth = 25 if max_depth is None else int(max_depth)
    split_threshold = 2 if split_threshold is None else int(split_threshold)
    c_trav = 1.0 if traversal_cost is None else float(traversal_cost)
    c_isect = 4.0 if intersection_cost is None else float(intersection_cost)
    b_lo, b_hi = lo.min(axis=0), hi.max(axis=0)
    pad = 1e-5 * np.maximum(b_hi - b_lo, 1e-6)
    b_lo, b_hi = b_lo - pad, b_hi + pad

    def area(e):                                    # (d-1)-dimensional measure of the box surface
        e = np.maximum(e, 1e-12)
        return float(np.sum(np.prod(e) / e))

    def make(idx, nlo, nhi, depth):
        m = idx.size
        if m == 0:
            return None
        if m <= split_threshold or depth >= max_depth:
            return T.KDLeaf([prims[i] for i in idx])
        ext = nhi - nlo
        parent_area = area(ext)
        best = (c_isect * m, None, None)
        for ax in np.argsort(-ext)[:min(d, 3)]:
            if ext[ax] <= 0:
                continue
            s_lo, s_hi = np.sort(lo[idx, ax]), np.sort(hi[idx, ax])
            cand = np.unique(np.concatenate([s_lo, s_hi]))
            cand = cand[(cand > nlo[ax]) & (cand < nhi[ax])]
            if cand.size == 0:
                continue
            if cand.size > 64:
                cand = cand[np.linspace(0, cand.size - 1, 64).astype(int)]
            n_left = np.searchsorted(s_lo, cand, side='left')         # boxes starting before the plane
            n_right = m - np.searchsorted(s_hi, cand, side='right')   # boxes ending after the plane
            e = ext.copy()
            for c, nl, nr in zip(cand, n_left, n_right):
                e[ax] = c - nlo[ax]
                al = area(e)
                e[ax] = nhi[ax] - c
                ar = area(e)
                cost = c_trav + c_isect * (al * nl + ar * nr) / parent_area
                if cost < best[0]:
                    best = (cost, int(ax), float(c))
        if best[1] is None:
            return T.KDLeaf([prims[i] for i in idx])
        ax, split = best[1], float(np.float32(best[2]))
        left_idx = idx[lo[idx, ax] < split]
        right_idx = idx[hi[idx, ax] > split]
        flat = idx[(lo[idx, ax] == split) & (hi[idx, ax] == split)]    # lying in the plane: keep on both sides
        left_idx = np.union1d(left_idx, flat)
        right_idx = np.union1d(right_idx, flat)
        if left_idx.size == m and right_idx.size == m:
            return T.KDLeaf([prims[i] for i in idx])
        l_hi, r_lo = nhi.copy(), nlo.copy()
        l_hi[ax], r_lo[ax] = split, split
        left = make(left_idx, nlo, l_hi, depth + 1)
        right = make(right_idx, r_lo, nhi, depth + 1)
        if left is None and right is None:
            return None
        return T.KDBranch(ax, split, left, right)

    root = make(np.arange(n), b_lo, b_hi, 0)
    boundary = T.AABB(d, T.Vector._wrap(b_lo.astype(np.float32)), T.Vector._wrap(b_hi.astype(np.float32)))
    return boundary, root
