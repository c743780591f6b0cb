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


def batched_soup(dim, n_batches, seed=7, batch=4):
    """Synthetic CompositeScene in `dim` dimensions made of `n_batches` 4-lane TriangleBatches (what the reference's SSE
    flavour stores in its leaves) plus a few single simplexes, over a tree from this repo's builder, with two
    materials (one reflective), a point light, a global light and shadows: exercises batch_test / simplex_single /
    shading / shadow rays / one reflection pass of the fixed-dimension kernels for which the reference's own fixtures
    (dimensions 3, 4, 6, 9) have no scene.  Not a reference golden: parity is against the oracle restatement."""
    from ntracer_b200 import bulk
    rng = np.random.RandomState(seed)
    n_single = 5
    n = n_batches * batch + n_single
    # A random (D-1)-simplex almost never meets the camera's 3-flat inside its barycentric range once D > 5 (SURVEY 8d,
    # C5), so each simplex is a triangle in the camera's 3-space (slightly below the flat in the extra axes) joined to
    # one far vertex along every extra axis: its slice with the 3-flat is close to the triangle itself.
    pts = np.zeros((n, dim, dim), np.float64)
    c = rng.uniform(-1, 1, (n, 1, 3))
    pts[:, :3, :3] = c + rng.uniform(-0.45, 0.45, (n, 3, 3))
    pts[:, :3, 3:] = -rng.uniform(0.01, 0.05, (n, 3, dim - 3))
    for k in range(3, dim):
        pts[:, k, :3] = pts[:, :3, :3].mean(axis=1) + rng.uniform(-0.05, 0.05, (n, 3))
        pts[:, k, 3:] = -rng.uniform(0.01, 0.05, (n, dim - 3))
        pts[:, k, k] = rng.uniform(0.5, 1.5, n)
    pts = pts.astype(np.float32)
    rec = bulk.simplex_records(pts)
    lo, hi = pts.min(axis=1), pts.max(axis=1)
    # items: batches of `batch` consecutive records first, then the singles
    ilo = [lo[k * batch:(k + 1) * batch].min(axis=0) for k in range(n_batches)] + [lo[n_batches * batch + k] for k in range(n_single)]
    ihi = [hi[k * batch:(k + 1) * batch].max(axis=0) for k in range(n_batches)] + [hi[n_batches * batch + k] for k in range(n_single)]
    refs_of_item = np.array([(1 << 30) | (k * batch) for k in range(n_batches)] +
                            [n_batches * batch + k for k in range(n_single)], dtype=np.uint32)
    nodes, item_idx, root, boundary = bulk.build_kdtree(np.array(ilo, np.float32), np.array(ihi, np.float32),
                                                        max_depth=8, split_threshold=3)
    # the reference keeps the batches of a leaf in front of its single primitives (tracer.hpp:1142-1150)
    nodes = nodes.copy()
    refs = refs_of_item[item_idx]
    for k in range(len(nodes)):
        if nodes[k, 0] & 0x80000000:
            a, m = int(nodes[k, 1]), int(nodes[k, 2])
            seg = refs[a:a + m]
            isb = (seg >> 30) == 1
            refs[a:a + m] = np.concatenate([seg[isb], seg[~isb]])
            nodes[k, 0] = 0x80000000 | int(isb.sum())
    mats = np.array([[1, 0.5, 0.5, 1, 1, 1, 1, 0, 1, 8],
                     [0.4, 0.7, 1.0, 1, 1, 1, 1, 0.35, 0.8, 12]], dtype=np.float32)
    light_pos = np.zeros(dim, np.float32); light_pos[1] = 4; light_pos[2] = -4
    gdir = np.zeros(dim, np.float32); gdir[1] = -1
    cam_origin = np.zeros(dim, np.float32); cam_origin[2] = -3.5
    return {
        'dim': np.int64(dim), 'kind': np.int64(1), 'batch_size': np.int64(batch), 'root': np.int64(root),
        'nodes': nodes, 'leaf_refs': refs.astype(np.uint32), 'simplex': rec,
        'simplex_mat': (rng.uniform(size=n) < 0.4).astype(np.int32),
        'solids': np.zeros((0, 1 + 2 * dim * dim + dim), np.float32), 'solid_mat': np.zeros(0, np.int32),
        'materials': mats, 'boundary': boundary,
        'params': np.array([0.8, 1, 1, 2, 1], dtype=np.float64),            # fov, shadows, camera_light, max_reflect_depth, bg axis
        'ambient': np.full(3, 0.05, np.float32), 'bg1': np.ones(3, np.float32), 'bg2': np.zeros(3, np.float32),
        'bg3': np.array([0, 1, 1], np.float32),
        'point_lights': np.concatenate([light_pos, [30, 30, 30]]).astype(np.float32)[None],
        'global_lights': np.concatenate([gdir, [0.4, 0.4, 0.4]]).astype(np.float32)[None],
        'cam_origin': cam_origin, 'cam_axes': np.eye(dim, dtype=np.float32),
    }


def fuzz_scene(dim, seed, max_batches=24):
    """Random mixed scene for differential tests: fixtures.batched_soup geometry with four materials (opaque,
    reflective, transparent, transparent + reflective) dealt at random, 0-3 rotated / scaled hypercubes and hyperspheres
    at the origin (referenced from every leaf), random shadows / camera light / reflection depth / background axis."""
    rng = np.random.RandomState(seed)
    sc = batched_soup(dim, int(rng.randint(min(3, max_batches), max_batches + 1)), seed=seed)
    n = sc['simplex'].shape[0]
    sc['materials'] = np.array([[1, 0.5, 0.5, 1, 1, 1, 1, 0, 1, 8], [0.4, 0.7, 1.0, 1, 1, 1, 1, 0.35, 0.8, 12],
                                [0.9, 0.9, 0.2, 1, 1, 1, 0.5, 0, 1, 8], [0.2, 0.9, 0.4, 1, 1, 1, 0.6, 0.25, 0.5, 4]], np.float32)
    sc['simplex_mat'] = rng.choice(4, size=n, p=[0.4, 0.25, 0.2, 0.15]).astype(np.int32)
    ns = int(rng.randint(0, 4))
    if ns:
        rows = []
        for _ in range(ns):
            typ = float(rng.choice([1, 2]))
            i, j = rng.choice(dim, 2, replace=False)
            th = rng.uniform(-1, 1)
            rot = np.eye(dim, dtype=np.float32)
            rot[i, i] = rot[j, j] = np.cos(th)
            rot[i, j], rot[j, i] = -np.sin(th), np.sin(th)
            ori = (rot * rng.uniform(0.3, 0.8)).astype(np.float32)
            # position 0 keeps orientation*position == position, where the reference is self-consistent
            rows.append(np.concatenate([[typ], ori.ravel(), np.linalg.inv(ori).astype(np.float32).ravel(), np.zeros(dim, np.float32)]))
        sc['solids'] = np.array(rows, np.float32)
        sc['solid_mat'] = rng.choice(4, size=ns).astype(np.int32)
        nodes, refs = sc['nodes'].copy(), []
        for k in range(len(nodes)):
            if nodes[k, 0] & 0x80000000:
                a, m = int(nodes[k, 1]), int(nodes[k, 2])
                seg = list(sc['leaf_refs'][a:a + m]) + [(2 << 30) | q for q in range(ns)]
                nodes[k, 1], nodes[k, 2] = len(refs), len(seg)
                refs += seg
        sc['nodes'], sc['leaf_refs'] = nodes, np.array(refs, np.uint32)
    sc['params'] = np.array([0.8, rng.randint(0, 2), rng.randint(0, 2), rng.randint(0, 4), rng.randint(0, dim)], np.float64)
    return sc


def stacked_layers(n_layers, opacity=0.5):
    """n_layers big transparent triangles stacked along the view axis in front of an opaque one (3-D): every central ray
    collects n_layers transparent hits -- beyond the 10 the reference preallocates and, from 17 on, beyond the 16 the
    kernels keep (ntr_counters.truncated_hit_lists)."""
    from ntracer_b200 import bulk
    pts = np.zeros((n_layers + 1, 3, 3), np.float32)
    for k in range(n_layers + 1):
        z = 0.2 * k
        pts[k] = [[-2, -2, z], [2, -2, z + 0.01 * k], [0, 2.5, z]]
    mats = np.array([[0.3, 0.6, 1.0, 1, 1, 1, opacity, 0, 1, 8], [1, 0.5, 0.2, 1, 1, 1, 1, 0, 1, 8]], np.float32)
    mid = np.zeros(n_layers + 1, np.int32)
    mid[-1] = 1
    sc = bulk.simplex_scene(pts, material_ids=mid, materials=mats, max_depth=2)
    sc['cam_origin'] = np.array([0, 0, -4], np.float32)
    return sc
