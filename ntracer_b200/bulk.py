"""Bulk scene construction through the native host-side builder (csrc/builder.cpp): what a script does with
TrianglePrototype + build_composite_scene, for sizes where one Python object per primitive is not an option
(BASELINE config 5: 1 M ten-dimensional simplexes).  Produces flat scene dicts for DeviceScene.

The arithmetic follows the reference (Triangle.from_points, src/tracer.hpp:442-462); the tree comes from this
repo's own builder, not the reference's (nearest hits do not depend on the tree)."""
import ctypes as C

import numpy as np

from . import _capi


def simplex_records(points):
    """points: float32 [n, D, D] (n simplexes, D vertices of D coordinates) -> records float32 [n, (D+1)*D+1]"""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, d, d2 = pts.shape
    if d != d2:
        raise ValueError('a simplex in D dimensions needs exactly D points')
    rec = np.zeros((n, (d + 1) * d + 1), dtype=np.float32)
    _capi.check(_capi.load().ntr_simplex_from_points(d, n, pts.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p)))
    return rec


def build_kdtree(lo, hi, max_depth=0, split_threshold=0, traversal_cost=-1.0, intersection_cost=-1.0, cull=None):
    """lo, hi: float32 [n, D] item bounds -> (nodes uint32 [m,4], item indices per leaf uint32 [k], root, boundary [2,D])

    cull = (item_first uint32 [n+1], s_lo [m, D], s_hi [m, D], records [m, (D+1)*D+1]): the simplexes behind the items
    (item i owns simplexes item_first[i] .. item_first[i+1]-1; bounds and record of each) -- an item is then listed only
    in the cells one of its simplexes can touch (ntr_build_kdtree_culled)."""
    lo = np.ascontiguousarray(lo, dtype=np.float32)
    hi = np.ascontiguousarray(hi, dtype=np.float32)
    n, d = lo.shape
    if hi.shape != lo.shape:
        raise ValueError('lo and hi must have the same shape')
    if n and not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi)) and np.all(lo <= hi)):
        raise ValueError('item bounds must be finite with lo <= hi on every axis')
    lib = _capi.load()
    nodes_p, refs_p = C.c_void_p(), C.c_void_p()
    n_nodes, n_refs, root = C.c_uint32(), C.c_uint32(), C.c_uint32()
    boundary = np.zeros((2, d), dtype=np.float32)
    if cull is not None:
        first = np.ascontiguousarray(cull[0], dtype=np.uint32)
        s_lo, s_hi, s_pl = (np.ascontiguousarray(a, dtype=np.float32) for a in cull[1:])
        m = s_lo.shape[0] if s_lo.ndim == 2 else 0
        if first.shape != (n + 1,) or s_lo.shape != (m, d) or s_hi.shape != (m, d) or s_pl.shape != (m, (d + 1) * d + 1):
            raise ValueError('cull = (item_first [n+1], s_lo [m,D], s_hi [m,D], records [m,(D+1)*D+1])')
        _capi.check(lib.ntr_build_kdtree_culled(d, n, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                                first.ctypes.data_as(C.c_void_p), m, s_lo.ctypes.data_as(C.c_void_p),
                                                s_hi.ctypes.data_as(C.c_void_p), s_pl.ctypes.data_as(C.c_void_p), int(max_depth),
                                                int(split_threshold), C.c_float(traversal_cost), C.c_float(intersection_cost),
                                                C.byref(nodes_p), C.byref(n_nodes), C.byref(refs_p), C.byref(n_refs), C.byref(root),
                                                boundary.ctypes.data_as(C.c_void_p)))
    else:
        _capi.check(lib.ntr_build_kdtree(d, n, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), int(max_depth),
                                         int(split_threshold), float(traversal_cost), float(intersection_cost),
                                         C.byref(nodes_p), C.byref(n_nodes), C.byref(refs_p), C.byref(n_refs), C.byref(root),
                                         boundary.ctypes.data_as(C.c_void_p)))
    try:
        nodes = np.ctypeslib.as_array(C.cast(nodes_p, C.POINTER(C.c_uint32)), shape=(max(n_nodes.value, 1) * 4,))[:n_nodes.value * 4].copy().reshape(-1, 4)
        refs = np.ctypeslib.as_array(C.cast(refs_p, C.POINTER(C.c_uint32)), shape=(max(n_refs.value, 1),))[:n_refs.value].copy()
    finally:
        lib.ntr_free(nodes_p)
        lib.ntr_free(refs_p)
    return nodes, refs, int(root.value), boundary


def group_items(lo, hi, group=4):
    """Permutation of the n items (bounds lo/hi: [n, D]) in which every consecutive run of `group` entries is one batch of
    spatially close items (ntr_group_items: what the reference's group_primitives does before its tree build)."""
    lo = np.ascontiguousarray(lo, dtype=np.float32)
    hi = np.ascontiguousarray(hi, dtype=np.float32)
    n, d = lo.shape
    order = np.zeros(n, dtype=np.uint32)
    _capi.check(_capi.load().ntr_group_items(d, n, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), int(group),
                                             order.ctypes.data_as(C.c_void_p)))
    return order


def batched_tree(lo, hi, batch=4, records=None, **tree_kw):
    """Groups n simplexes into batches of `batch` and builds the tree over the ITEMS (batches + left-over singles), the
    way build_kdtree does in the reference's SIMD builds (src/tracer.hpp:2431-2455).
    -> (order [n]: new position -> old simplex index, nodes, leaf refs ((1<<30)|first record for batches), root, boundary)"""
    n = lo.shape[0]
    order = group_items(lo, hi, batch) if batch > 1 else np.arange(n, dtype=np.uint32)
    lo, hi = lo[order], hi[order]
    nb = n // batch if batch > 1 else 0
    ilo = np.concatenate([lo[:nb * batch].reshape(nb, batch, -1).min(axis=1), lo[nb * batch:]]) if nb else lo
    ihi = np.concatenate([hi[:nb * batch].reshape(nb, batch, -1).max(axis=1), hi[nb * batch:]]) if nb else hi
    item_ref = np.concatenate([(1 << 30) | (np.arange(nb, dtype=np.uint32) * batch),
                               np.arange(nb * batch, n, dtype=np.uint32)]).astype(np.uint32)
    cull = None
    if records is not None:
        # (records: [n, (D+1)*D+1] in the ORIGINAL order) item k < nb owns simplexes k*batch .. (k+1)*batch-1
        first = np.concatenate([np.arange(nb + 1, dtype=np.uint32) * batch, nb * batch + 1 + np.arange(n - nb * batch, dtype=np.uint32)])
        cull = (first, lo, hi, np.asarray(records, np.float32)[order])
    nodes, items, root, boundary = build_kdtree(ilo, ihi, cull=cull, **tree_kw)
    nodes = nodes.copy()
    refs = item_ref[items]
    # the reference keeps the batches of a leaf in front of its single primitives (tracer.hpp:1142-1150) and stores
    # their number in the leaf
    for k in np.nonzero(nodes[:, 0] & 0x80000000)[0]:
        a, m = int(nodes[k, 1]), int(nodes[k, 2])
        seg = refs[a:a + m]
        isb = (seg >> 30) == 1
        if isb.any() and not isb.all():
            refs[a:a + m] = np.concatenate([seg[isb], seg[~isb]])
        nodes[k, 0] = 0x80000000 | int(isb.sum())
    return order, nodes, refs, root, boundary


def simplex_scene(points, material_ids=None, materials=None, batch=1, cull=False, **tree_kw):
    """Flat CompositeScene dict for n simplexes given by their vertices (float32 [n, D, D]).  batch = 4 packs them into
    4-lane batch items first (the layout the tuned batch test of the kernels works on).  cull = True: the tree lists a
    simplex only in the cells it can touch (ntr_build_kdtree_culled) instead of every cell its box overlaps."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, d, _ = pts.shape
    rec = simplex_records(pts)
    lo, hi = pts.min(axis=1), pts.max(axis=1)
    if material_ids is None:
        material_ids = np.zeros(n, dtype=np.int32)
    material_ids = np.ascontiguousarray(material_ids, dtype=np.int32)
    if batch > 1:
        order, nodes, refs, root, boundary = batched_tree(lo, hi, batch, records=rec if cull else None, **tree_kw)
        rec, material_ids = rec[order], material_ids[order]
    else:
        nodes, refs, root, boundary = build_kdtree(lo, hi, cull=(np.arange(n + 1, dtype=np.uint32), lo, hi, rec) if cull else None, **tree_kw)
    if materials is None:
        materials = np.array([[1, 0.5, 0.5, 1, 1, 1, 1, 0, 1, 8]], dtype=np.float32)      # Material((1,0.5,0.5))
    cam_axes = np.eye(d, dtype=np.float32)
    return {
        'dim': np.int64(d), 'kind': np.int64(1), 'batch_size': np.int64(batch), 'root': np.int64(root),
        'nodes': nodes, 'leaf_refs': refs.astype(np.uint32),           # single simplex: item index == record index (type 0)
        'simplex': np.ascontiguousarray(rec), 'simplex_mat': material_ids,
        'solids': np.zeros((0, 1 + 2 * d * d + d), np.float32), 'solid_mat': np.zeros(0, np.int32),
        'materials': np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 10), 'boundary': boundary,
        'params': np.array([0.8, 0, 1, 4, 1], dtype=np.float64),
        'ambient': np.zeros(3, np.float32), 'bg1': np.ones(3, np.float32), 'bg2': np.zeros(3, np.float32),
        'bg3': np.array([0, 1, 1], np.float32),
        'point_lights': np.zeros((0, d + 3), np.float32), 'global_lights': np.zeros((0, d + 3), np.float32),
        'cam_origin': np.zeros(d, np.float32), 'cam_axes': cam_axes,
    }


def soup(dim, n, seed=1234, spread=0.05, thin=0.02):
    """The synthetic simplex soup of BASELINE config 5: centres U(-1,1)^3 x U(-thin,thin)^(D-3),
    vertices centre + U(-spread,spread)^D."""
    rng = np.random.RandomState(seed)
    c = np.concatenate([rng.uniform(-1, 1, (n, 3)), rng.uniform(-thin, thin, (n, dim - 3))], axis=1).astype(np.float32)
    pts = c[:, None, :] + rng.uniform(-spread, spread, (n, dim, dim)).astype(np.float32)
    return pts.astype(np.float32)
