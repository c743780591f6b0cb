"""The native host-side builder (csrc/builder.cpp: SURVEY section 8f rows 1-2): Triangle.from_points in bulk and the
k-d tree.  CPU only."""
import numpy as np
import pytest

from ntracer_b200 import NTracer, Material, bulk, _capi
from tests import oracle_lib as ol


@pytest.mark.parametrize('dim', [3, 4, 5, 7, 10])
def test_simplex_records_match_from_points(dim):
    pts = bulk.soup(dim, 40, seed=dim, spread=0.5)
    rec = bulk.simplex_records(pts)
    nt = NTracer(dim)
    for k in (0, 17, 39):
        t = nt.Triangle.from_points([nt.Vector(*p) for p in pts[k]], Material((1, 1, 1)))
        row = t._row()
        assert np.abs(rec[k] - row).max() <= 2e-6 * np.abs(row).max()
    # defining property (reference test_to_from_points, any dimension): edge_normal_i . (p_{j+1} - p_0) = -delta_ij
    D = dim
    E = rec[:, 2 * D + 1:].reshape(-1, D - 1, D)
    V = pts[:, 1:, :] - pts[:, :1, :]
    prod = np.einsum('nid,njd->nij', E.astype(np.float64), V.astype(np.float64))
    assert np.abs(prod + np.eye(D - 1)).max() < 2e-3
    assert np.abs(np.einsum('nd,njd->nj', rec[:, :D].astype(np.float64), V.astype(np.float64))).max() < 1e-4 * np.abs(rec[:, :D]).max()


def walk(nodes, refs, root, lo, hi, fn):
    stack = [(root, lo.copy(), hi.copy())]
    while stack:
        n, a, b = stack.pop()
        if n == _capi.NULL_NODE:
            fn(None, a, b)
            continue
        meta, w1, w2, w3 = (int(x) for x in nodes[n])
        if meta & _capi.LEAF_FLAG:
            fn(refs[w1:w1 + w2], a, b)
        else:
            split = float(np.array([w1], np.uint32).view(np.float32)[0])
            lb, ra = b.copy(), a.copy()
            lb[meta], ra[meta] = split, split
            stack.append((w2, a, lb))
            stack.append((w3, ra, b))


@pytest.mark.parametrize('dim,n', [(3, 500), (5, 2000), (10, 3000)])
def test_tree_is_complete(dim, n):
    """Every item is listed in every leaf cell its box overlaps with positive measure, and empty (null) cells
    overlap no item: then any ray finds every primitive it can hit."""
    rng = np.random.RandomState(n)
    c = rng.uniform(-1, 1, (n, dim))
    e = rng.uniform(0.01, 0.2, (n, dim))
    lo, hi = (c - e).astype(np.float32), (c + e).astype(np.float32)
    nodes, refs, root, boundary = bulk.build_kdtree(lo, hi)
    assert np.all(boundary[0] <= lo.min(axis=0)) and np.all(boundary[1] >= hi.max(axis=0))
    seen = np.zeros(n, bool)
    checked = [0]

    def fn(items, a, b):
        inside = np.all((lo < b) & (hi > a), axis=1)           # overlap with positive measure
        if items is None:
            assert not inside.any()
            return
        members = np.zeros(n, bool)
        members[items] = True
        seen[items] = True
        assert not np.any(inside & ~members)
        checked[0] += 1

    walk(nodes, refs, root, boundary[0].astype(np.float64), boundary[1].astype(np.float64), fn)
    assert seen.all() and checked[0] > 1
    depth = _tree_depth(nodes, root)
    assert depth <= 62


def _tree_depth(nodes, root):
    best, stack = 0, [(root, 1)]
    while stack:
        n, d = stack.pop()
        if n == _capi.NULL_NODE:
            continue
        best = max(best, d)
        meta, w1, w2, w3 = (int(x) for x in nodes[n])
        if not meta & _capi.LEAF_FLAG:
            stack += [(w2, d + 1), (w3, d + 1)]
    return best


def test_bulk_scene_equals_single_leaf_scene_under_the_oracle():
    """config 5 in miniature: the images must not depend on which tree is used (no shadows)."""
    dim, n = 10, 1500
    pts = bulk.soup(dim, n, spread=0.5)
    sc = bulk.simplex_scene(pts)
    sc['cam_origin'] = np.array([0, 0, -3] + [0] * (dim - 3), np.float32)
    one = dict(sc)
    one['nodes'] = np.array([[_capi.LEAF_FLAG, 0, n, 0]], np.uint32)
    one['leaf_refs'] = np.arange(n, dtype=np.uint32)
    one['root'] = np.int64(0)
    a = ol.render_float(sc, 48, 27)
    b = ol.render_float(one, 48, 27)
    ia, da = ol.primary_hit_ids(sc, 48, 27)
    ib, db = ol.primary_hit_ids(one, 48, 27)
    assert (ia >= 0).mean() > 0.02
    assert np.mean(ia == ib) >= 0.999
    assert np.abs(a - b).max() < 1e-4


def test_builder_edge_cases_and_input_validation():
    import ctypes as C
    from ntracer_b200 import _capi
    one = np.array([[0, 0, 0]], np.float32)
    nodes, refs, root, bnd = bulk.build_kdtree(one, one + 1)
    assert nodes.shape == (1, 4) and refs.tolist() == [0] and root == 0 and np.all(bnd[0] < 0) and np.all(bnd[1] > 1)
    nodes, refs, root, bnd = bulk.build_kdtree(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert nodes.shape == (0, 4) and refs.size == 0 and root == 0xFFFFFFFF            # an empty scene: no root
    flat_lo = np.array([[0, 0, 0], [2, 0, 0]], np.float32)                            # zero extent on one axis
    nodes, refs, root, bnd = bulk.build_kdtree(flat_lo, flat_lo + np.array([1, 1, 0], np.float32))
    assert sorted(refs.tolist()) == [0, 1]
    for bad_lo, bad_hi in (([[0, 0, np.nan]], [[1, 1, 1]]), ([[1, 1, 1]], [[0, 0, 0]]), ([[0, 0, 0]], [[1, np.inf, 1]])):
        with pytest.raises(ValueError):
            bulk.build_kdtree(np.array(bad_lo, np.float32), np.array(bad_hi, np.float32))
    # the C entry point refuses the same input on its own (NTR_ERR_VALUE), whatever the caller checked
    lib = _capi.load()
    lo, hi = np.array([[1, 1, 1]], np.float32), np.array([[0, 0, 0]], np.float32)
    p1, p2, a, b, c = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    bnd = np.zeros((2, 3), np.float32)
    rc = lib.ntr_build_kdtree(3, 1, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), 0, 0, -1.0, -1.0,
                              C.byref(p1), C.byref(a), C.byref(p2), C.byref(b), C.byref(c), bnd.ctypes.data_as(C.c_void_p))
    assert rc == _capi.NTR_ERR_VALUE


def test_group_items_is_a_permutation_of_compact_groups():
    """ntr_group_items (batch grouping ahead of the tree build, the reference's group_primitives): a permutation in which
    consecutive runs of 4 are spatially close -- far tighter than the same items grouped in input order."""
    from ntracer_b200 import bulk
    rng = np.random.RandomState(5)
    for dim, n in ((3, 1003), (4, 4096), (7, 50), (10, 3)):
        c = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
        lo, hi = c - 0.01, c + 0.01
        order = bulk.group_items(lo, hi, 4)
        assert sorted(order.tolist()) == list(range(n))
        if n >= 50:
            nb = n // 4

            def spread(idx):
                g = c[idx[:nb * 4]].reshape(nb, 4, dim)
                return float((g.max(axis=1) - g.min(axis=1)).sum(axis=1).mean())
            assert spread(order) < (0.5 if n >= 1000 else 0.9) * spread(np.arange(n))      # (few points in 7-D: little to gain)
    with pytest.raises(ValueError):
        bulk.group_items(np.zeros((4, 2), np.float32), np.ones((4, 2), np.float32), 4)


def test_batched_scene_renders_like_the_single_simplex_scene():
    """bulk.simplex_scene(batch=4): records reordered into 4-lane batch items, tree over the items, batches first in every
    leaf -- the oracle must draw the same picture as for the single-simplex scene of the same soup."""
    from ntracer_b200 import bulk
    from tests import oracle_lib as ol
    pts = bulk.soup(4, 2003, thin=0.3, spread=0.2)
    a = bulk.simplex_scene(pts, max_depth=14)
    b = bulk.simplex_scene(pts, batch=4, max_depth=14)
    for sc in (a, b):
        sc['cam_origin'] = np.array([0, 0, -3, 0], np.float32)
    assert int(b['batch_size']) == 4 and int((b['leaf_refs'] >> 30 == 1).sum()) > 0
    n_batches = (b['nodes'][:, 0] & 0x7FFFFFFF)[(b['nodes'][:, 0] >> 31) == 1]
    for k in np.nonzero((b['nodes'][:, 0] >> 31) == 1)[0][:200]:         # batches come first in a leaf
        first, m, nb = int(b['nodes'][k, 1]), int(b['nodes'][k, 2]), int(b['nodes'][k, 0] & 0x7FFFFFFF)
        kinds = b['leaf_refs'][first:first + m] >> 30
        assert np.all(kinds[:nb] == 1) and np.all(kinds[nb:] == 0)
    assert np.array_equal(ol.render_float(a, 96, 54), ol.render_float(b, 96, 54))
    ids_a, _ = ol.primary_hit_ids(a, 96, 54)
    assert (ids_a >= 0).mean() > 0.2
