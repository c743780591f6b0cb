"""The native host-side builder (csrc/builder.cpp: SURVEY section 8f rows 1-2): Triangle.from_points in bulk and the
k-d tree.  CPU only."""
import numpy as np
import pytest

from ntracer_b200 import NTracer, Material, bulk, _capi
from tests import emul_lib as el
from tests import fixtures as fx
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


def _leaf_of(nodes, root, x):
    k = root
    while True:
        meta, a, b, c = (int(v) for v in nodes[k])
        if meta & 0x80000000:
            return k
        split = float(np.array([a], np.uint32).view(np.float32)[0])
        k = b if x[meta] < split else c
        if k == 0xFFFFFFFF:
            return None


@pytest.mark.parametrize('dim,n', [(3, 400), (4, 600), (6, 300), (10, 200)])
def test_culled_tree_never_drops_a_simplex_from_a_cell_it_touches(dim, n):
    """ntr_build_kdtree_culled lists an item only in the cells one of its simplexes can touch (separating axes: bounds,
    face normal, facet directions -- what the reference's builder does with its exact overlap tests,
    src/tracer.hpp:1465-1675).  Conservative: every point OF a simplex lies in a leaf that lists the simplex (random
    barycentric samples, vertices and facet centres included); and, for the flat simplexes of 3 and 4 dimensions, much
    smaller than the tree over bounding boxes."""
    from ntracer_b200 import bulk
    rng = np.random.RandomState(dim)
    pts = bulk.soup(dim, n, seed=dim, spread=0.3)
    rec = bulk.simplex_records(pts)
    lo, hi = pts.min(axis=1), pts.max(axis=1)
    plain = bulk.build_kdtree(lo, hi, max_depth=18)
    nodes, refs, root, boundary = bulk.build_kdtree(lo, hi, max_depth=18, cull=(np.arange(n + 1, dtype=np.uint32), lo, hi, rec))
    if dim <= 4:
        assert len(refs) < 0.85 * len(plain[1])
    lam = rng.dirichlet(np.ones(dim) * 0.5, size=(n, 24)).astype(np.float64)          # 24 samples per simplex, corners favoured
    lam[:, 0] = np.eye(dim)[rng.randint(0, dim, n)]                                    # a vertex
    lam[:, 1] = (1 - np.eye(dim)[rng.randint(0, dim, n)]) / (dim - 1)                  # a facet centre
    missing = 0
    for s in range(n):
        for x in lam[s] @ pts[s].astype(np.float64):
            leaf = _leaf_of(nodes, root, x)
            if leaf is None or s not in refs[int(nodes[leaf, 1]):int(nodes[leaf, 1]) + int(nodes[leaf, 2])]:
                # a sample within rounding of a split plane may sit on the other side of it: the neighbour must list it then
                near = False
                k = root
                while k != 0xFFFFFFFF and not (int(nodes[k, 0]) & 0x80000000):
                    split = float(np.array([int(nodes[k, 1])], np.uint32).view(np.float32)[0])
                    if abs(x[int(nodes[k, 0])] - split) < 1e-5:
                        near = True
                    k = int(nodes[k, 2]) if x[int(nodes[k, 0])] < split else int(nodes[k, 3])
                missing += not near
    assert missing == 0


def test_culled_tree_renders_the_same_picture():
    """Scenes built with and without the culling, single simplexes and 4-lane batches: the oracle's hit ids and distances
    are identical, the lit picture is the same, shadows included (an occluder is found in whichever cell the shadow ray
    meets it), and the culled tree costs fewer simplex tests."""
    from ntracer_b200 import bulk
    w, h = 96, 54
    for dim, n, batch in ((4, 500, 1), (4, 500, 4), (7, 240, 4)):
        pts = bulk.soup(dim, n, seed=3 + dim, spread=0.3)
        out = []
        for cull in (False, True):
            sc = bulk.simplex_scene(pts, batch=batch, cull=cull, max_depth=16)
            sc['cam_origin'] = np.array([0, 0, -3] + [0] * (dim - 3), np.float32)
            sc['point_lights'] = np.array([[2, 3, -4] + [0] * (dim - 3) + [1, 1, 1]], np.float32)
            sc['params'] = np.array([0.8, 1, 1, 4, 1], dtype=np.float64)                   # shadows on
            img, cnt = ol.render_float(sc, w, h, with_counters=True)
            ids, dist = ol.primary_hit_ids(sc, w, h)
            out.append((sc, img, cnt, ids, dist))
        (s0, i0, c0, id0, d0), (s1, i1, c1, id1, d1) = out
        assert (id0 >= 0).sum() > 50
        assert np.array_equal(id0 >= 0, id1 >= 0) and np.array_equal(d0, d1)
        bad, mx = fx.lsb_stats(i1, i0)
        assert bad <= 0.001, (dim, batch, bad, mx)
        assert c1['simplex_tests'] < c0['simplex_tests'], (dim, batch, c1['simplex_tests'], c0['simplex_tests'])
        e, _ = el.render(s1, w, h)
        assert np.abs(e - i1).max() <= 2e-6                                              # the device code on the culled tree


def test_rebuilt_120_cell_costs_what_the_reference_tree_costs():
    """The {5,3,3} fixture rebuilt from its own simplexes through this repo's grouping and builder (what
    build_composite_scene does) against the tree the reference built: same hits, the shadowed picture within the parity
    bound (the only tree-dependent part), and no more simplex tests per frame than on the
    reference's tree (2.2 x with bounding boxes alone)."""
    from ntracer_b200 import bulk
    sc, g = fx.load('cell120')
    sc = fx.variant(sc, g, 'shadows')
    D, rec = int(sc['dim']), sc['simplex']
    n = rec.shape[0]
    p1 = rec[:, D + 1:2 * D + 1].astype(np.float64)
    E = rec[:, 2 * D + 1:2 * D + 1 + (D - 1) * D].reshape(n, D - 1, D).astype(np.float64)
    pts = np.zeros((n, D, D))
    pts[:, 0] = p1
    for k in range(n):
        pts[k, 1:] = p1[k] - np.linalg.pinv(E[k]).T        # edge_normal_i . (p1 - p_j) = delta_ij (tracer.hpp:454-461)
    pts = pts.astype(np.float32)
    assert np.abs(bulk.simplex_records(pts) - rec).max() <= 1e-4
    b = bulk.simplex_scene(pts, material_ids=sc['simplex_mat'], materials=sc['materials'], batch=4, cull=True)
    for k in ('params', 'ambient', 'bg1', 'bg2', 'bg3', 'point_lights', 'global_lights', 'cam_origin', 'cam_axes'):
        b[k] = sc[k]
    w, h = 240, 135
    ia, ca = ol.render_float(sc, w, h, with_counters=True)
    ib, cb = ol.render_float(b, w, h, with_counters=True)
    bad, mx = fx.lsb_stats(ib, ia)
    assert bad <= 0.001, (bad, mx)
    assert cb['shadow_rays'] == ca['shadow_rays']
    assert cb['simplex_tests'] <= 1.05 * ca['simplex_tests'] and cb['node_steps'] <= 1.05 * ca['node_steps']
