"""TEST INFRASTRUCTURE ONLY -- bridge between the compiled reference (oracle/_ref, built by
oracle/build_ref.sh from the unmodified Rouslan/NTracer sources) and this repo's flat scene files.

Nothing under ntracer_b200/ imports this module.  It is used by
  * tests/golden/make_fixtures.py  (run in the build container, where /root/reference exists) to
    export reference-built scenes (k-d tree included) and golden outputs into tests/golden/, and
  * bench.py's cpu_baseline / --impl reference leg and the `-m gpu` parity tests, which rebuild the
    *same* scene (same tree, same leaf order) inside the reference through its public constructors
    and time / query the reference's own CPU renderer.

Flat scene format (dict of numpy arrays, stored as .npz) -- see DESIGN.md "scene file":
  dim, kind (0 = BoxScene, 1 = CompositeScene), batch_size
  nodes      uint32 [n,4]   branch: {axis, float32 bits of split, left, right} (0xFFFFFFFF = null child,
                            reference src/tracer.hpp:813-817); leaf: {0x80000000|n_batches, first_ref, n_items, 0}
  leaf_refs  uint32 [m]     (type<<30)|index ; type 0 = single simplex, 1 = batch (index = first simplex of
                            `batch_size` consecutive ones), 2 = solid.  Order = the reference's leaf order.
  simplex    float32 [ns,(D+1)*D+1]  face_normal[D], d, p1[D], edge_normals[D-1][D]   (tracer.hpp:392-401)
  simplex_mat int32 [ns];  solids float32 [nsol, 1+2*D*D+D] = type, orientation, inv_orientation, position
  solid_mat  int32 [nsol]; materials float32 [nm,10] = color, specular, opacity, reflectivity,
                            specular_intensity, specular_exp  (render.hpp:56-73)
  root int64, boundary float32 [2,D], params (fov, shadows, camera_light, max_reflect_depth, bg_gradient_axis),
  ambient/bg1/bg2/bg3 float32 [3], point_lights/global_lights float32 [n, D+3], cam_origin [D], cam_axes [D,D]
"""
import os
import sys

import numpy as np

NULL = 0xFFFFFFFF
LEAF_FLAG = 0x80000000
REF_SIMPLEX, REF_BATCH, REF_SOLID = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')


def have_reference():
    return os.path.isdir(os.path.join(REF_DIR, 'ntracer'))


def load_reference():
    """Import the compiled reference package from oracle/_ref and return the `ntracer` module."""
    if not have_reference():
        raise RuntimeError('oracle/_ref is missing: run oracle/build_ref.sh in the build container')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import ntracer
    return ntracer


def _f32(v):
    return np.array(list(v), dtype=np.float32)


def _mat_row(m):
    c, s = m.color, m.specular
    return np.array([c.r, c.g, c.b, s.r, s.g, s.b, m.opacity, m.reflectivity,
                     m.specular_intensity, m.specular_exp], dtype=np.float32)


def export_scene(nt, scene):
    """Walk a reference BoxScene/CompositeScene through its public Python API into the flat format."""
    dim = nt.dimension
    cam = scene.get_camera()
    out = {
        'dim': np.int64(dim),
        'batch_size': np.int64(nt.BATCH_SIZE),
        'cam_origin': _f32(cam.origin),
        'cam_axes': np.stack([_f32(cam.axes[i]) for i in range(dim)]),
    }
    if isinstance(scene, nt.base.BoxScene):
        out['kind'] = np.int64(0)
        out['params'] = np.array([scene.fov, 0, 0, 0, 0], dtype=np.float64)
        return out

    out['kind'] = np.int64(1)
    mats, mat_ids = [], {}
    simplex, simplex_mat = [], []
    solids, solid_mat = [], []
    item_ref = {}      # id(python primitive object) -> leaf ref value
    keep = []          # keep python objects alive so that id() stays unique
    nodes, refs = [], []

    def mat_id(m):
        k = id(m)
        if k not in mat_ids:
            keep.append(m)
            mat_ids[k] = len(mats)
            mats.append(_mat_row(m))
        return mat_ids[k]

    def add_triangle(t):
        row = np.concatenate([_f32(t.face_normal), np.array([t.d], dtype=np.float32), _f32(t.p1)] +
                             [_f32(e) for e in t.edge_normals])
        simplex.append(row)
        simplex_mat.append(mat_id(t.material))
        return len(simplex) - 1

    def ref_of(item):
        k = id(item)
        if k in item_ref:
            return item_ref[k]
        keep.append(item)
        if isinstance(item, nt.base.TriangleBatch):
            first = None
            for lane in range(nt.BATCH_SIZE):
                i = add_triangle(item[lane])
                if first is None:
                    first = i
            r = (REF_BATCH << 30) | first
        elif isinstance(item, nt.base.Triangle):
            r = (REF_SIMPLEX << 30) | add_triangle(item)
        elif isinstance(item, nt.base.Solid):
            row = np.concatenate([np.array([item.type], dtype=np.float32),
                                  _f32(item.orientation.values), _f32(item.inv_orientation.values),
                                  _f32(item.position)])
            solids.append(row)
            solid_mat.append(mat_id(item.material))
            r = (REF_SOLID << 30) | (len(solids) - 1)
        else:
            raise TypeError('unknown primitive type %r' % type(item))
        item_ref[k] = r
        return r

    def walk(node):
        if node is None:
            return NULL
        idx = len(nodes)
        nodes.append(None)
        if isinstance(node, nt.base.KDLeaf):
            items = [node[i] for i in range(len(node))]
            first = len(refs)
            nb = 0
            for it in items:
                r = ref_of(it)
                if (r >> 30) == REF_BATCH:
                    nb += 1
                refs.append(r)
            nodes[idx] = (LEAF_FLAG | nb, first, len(items), 0)
        else:
            split = np.array([node.split], dtype=np.float32).view(np.uint32)[0]
            left = walk(node.left)
            right = walk(node.right)
            nodes[idx] = (node.axis, int(split), left, right)
        return idx

    sys.setrecursionlimit(10000)
    root = walk(scene.root)
    stride = (dim + 1) * dim + 1
    out.update({
        'root': np.int64(root),
        'nodes': np.array(nodes, dtype=np.uint32).reshape(-1, 4),
        'leaf_refs': np.array(refs, dtype=np.uint32),
        'simplex': np.array(simplex, dtype=np.float32).reshape(-1, stride),
        'simplex_mat': np.array(simplex_mat, dtype=np.int32),
        'solids': np.array(solids, dtype=np.float32).reshape(-1, 1 + 2 * dim * dim + dim),
        'solid_mat': np.array(solid_mat, dtype=np.int32),
        'materials': np.array(mats, dtype=np.float32).reshape(-1, 10),
        'boundary': np.stack([_f32(scene.boundary.start), _f32(scene.boundary.end)]),
        'params': np.array([scene.fov, scene.shadows, scene.camera_light, scene.max_reflect_depth,
                            scene.bg_gradient_axis], dtype=np.float64),
        'ambient': _f32(scene.ambient_color), 'bg1': _f32(scene.bg1), 'bg2': _f32(scene.bg2),
        'bg3': _f32(scene.bg3),
        'point_lights': np.array([list(l.position) + list(l.color) for l in scene.point_lights],
                                 dtype=np.float32).reshape(-1, dim + 3),
        'global_lights': np.array([list(l.direction) + list(l.color) for l in scene.global_lights],
                                  dtype=np.float32).reshape(-1, dim + 3),
    })
    # id(python object) -> flat primitive id, for mapping RayIntersection.primitive back (make_fixtures)
    out['_item_ref'] = item_ref
    out['_keep'] = keep
    return out


def strip_private(sc):
    return {k: v for k, v in sc.items() if not k.startswith('_')}


def save_scene(path, sc):
    np.savez_compressed(path, **strip_private(sc))


def load_scene(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def flat_prim_id(sc, ref, lane):
    """Flat primitive id used by ntr_primary_hit_ids: simplex index (batch lanes are consecutive),
    solids follow the simplexes."""
    kind, idx = ref >> 30, ref & 0x3FFFFFFF
    if kind == REF_BATCH:
        return idx + lane
    if kind == REF_SIMPLEX:
        return idx
    return int(sc['simplex'].shape[0]) + idx


def import_scene(sc, force_generic=False):
    """Rebuild the flat scene inside the reference via its public constructors.  Returns
    (nt, scene, prim_objects) where prim_objects maps leaf ref value -> reference primitive object."""
    ntracer = load_reference()
    dim = int(sc['dim'])
    nt = ntracer.NTracer(dim, force_generic)
    cam = nt.Camera()
    cam.origin = nt.Vector(*[float(x) for x in sc['cam_origin']])
    for i in range(dim):
        cam.axes[i] = nt.Vector(*[float(x) for x in sc['cam_axes'][i]])
    fov = float(sc['params'][0])
    if int(sc['kind']) == 0:
        scene = nt.BoxScene()
        scene.set_camera(cam)
        scene.set_fov(fov)
        return nt, scene, {}

    if int(sc['batch_size']) != nt.BATCH_SIZE and np.any((sc['leaf_refs'] >> 30) == REF_BATCH):
        raise RuntimeError('scene file was exported with BATCH_SIZE=%d but this reference build has %d'
                           % (int(sc['batch_size']), nt.BATCH_SIZE))
    B = nt.BATCH_SIZE
    mats = []
    for r in sc['materials']:
        r = [float(x) for x in r]
        mats.append(ntracer.Material((r[0], r[1], r[2]), r[6], r[7], r[8], r[9], (r[3], r[4], r[5])))

    def vec(a):
        return nt.Vector(*[float(x) for x in a])

    tri_cache = {}

    def triangle(i):
        if i not in tri_cache:
            row = sc['simplex'][i]
            fn = row[0:dim]
            p1 = row[dim + 1:2 * dim + 1]
            edges = row[2 * dim + 1:].reshape(dim - 1, dim)
            tri_cache[i] = nt.Triangle(vec(p1), vec(fn), [vec(e) for e in edges], mats[int(sc['simplex_mat'][i])])
        return tri_cache[i]

    prims = {}

    def prim(ref):
        ref = int(ref)
        if ref in prims:
            return prims[ref]
        kind, idx = ref >> 30, ref & 0x3FFFFFFF
        if kind == REF_BATCH:
            p = nt.TriangleBatch([triangle(idx + l) for l in range(B)])
        elif kind == REF_SIMPLEX:
            p = triangle(idx)
        else:
            row = sc['solids'][idx]
            o = nt.Matrix(*[float(x) for x in row[1:1 + dim * dim]])
            pos = vec(row[1 + 2 * dim * dim:])
            p = nt.Solid(int(row[0]), pos, o, mats[int(sc['solid_mat'][idx])])
        prims[ref] = p
        return p

    nodes = sc['nodes']
    refs = sc['leaf_refs']

    def build(i):
        if i == NULL:
            return None
        meta, a, b, c = [int(x) for x in nodes[i]]
        if meta & LEAF_FLAG:
            return nt.KDLeaf([prim(r) for r in refs[a:a + b]])
        split = float(np.array([a], dtype=np.uint32).view(np.float32)[0])
        return nt.KDBranch(meta, split, build(b), build(c))

    sys.setrecursionlimit(10000)
    root = build(int(sc['root']))
    scene = nt.CompositeScene(nt.AABB(vec(sc['boundary'][0]), vec(sc['boundary'][1])), root)
    scene.set_camera(cam)
    scene.set_fov(fov)
    scene.set_shadows(bool(sc['params'][1]))
    scene.set_camera_light(bool(sc['params'][2]))
    scene.set_max_reflect_depth(int(sc['params'][3]))
    scene.set_ambient_color(tuple(float(x) for x in sc['ambient']))
    scene.set_background(tuple(float(x) for x in sc['bg1']), tuple(float(x) for x in sc['bg2']),
                         tuple(float(x) for x in sc['bg3']), int(sc['params'][4]))
    for l in sc['point_lights']:
        scene.add_light(nt.PointLight(vec(l[:dim]), tuple(float(x) for x in l[dim:])))
    for l in sc['global_lights']:
        scene.add_light(nt.GlobalLight(vec(l[:dim]), tuple(float(x) for x in l[dim:])))
    return nt, scene, prims


def make_immortal(objs):
    """Workaround for a data race in the UNMODIFIED reference, applied from the outside: kd_leaf::occludes copies
    a py::object per leaf item (src/tracer.hpp:1094 `auto item = this->items()[i];`), i.e. Py_INCREF/Py_DECREF on the
    primitive from every render thread without the GIL.  The non-atomic updates get lost, the count eventually hits
    zero and a primitive is freed under the renderer (ASAN: heap-use-after-free in a worker of BlockingRenderer;
    seen as SIGSEGV / 'corrupted double-linked list' at 1280x720 and above with shadows on).  CPython 3.12 skips
    reference counting for immortal objects, so marking the primitives immortal removes the racy writes without
    touching the reference.  (It also removes cache-line ping-pong: if anything this favours the CPU baseline.)"""
    import ctypes
    for o in objs:
        ctypes.c_uint32.from_address(id(o)).value = 0xFFFFFFFF     # _Py_IMMORTAL_REFCNT (64-bit builds)


def polytope_scene(schlafli, cam_dist=4.0):
    """Run the reference's own scripts/polytope.py geometry + tree build for a Schlafli symbol
    (e.g. '5 3 3', '5/2 3 3') and return (nt, scene, camera).  pygame is stubbed (it is UI only) and
    fractions.gcd (removed in Python 3.9) is aliased, exactly as SURVEY.md section 8(c) records."""
    import fractions
    import math
    import types
    load_reference()
    if not hasattr(fractions, 'gcd'):
        fractions.gcd = math.gcd
    sys.modules.setdefault('pygame', types.ModuleType('pygame'))
    sys.modules.setdefault('ntracer.pygame_render', types.ModuleType('ntracer.pygame_render'))
    sys.modules['ntracer.pygame_render'].PygameRenderer = object
    path = os.path.join(REF_DIR, 'scripts', 'polytope.py')
    src = open(path).read()
    src = src[:src.index('if args.output is not None:')]
    argv = sys.argv
    sys.argv = ['polytope.py'] + schlafli.split() + ['-d', str(cam_dist)]
    g = {'__name__': 'polytope_ref'}
    hook = sys.excepthook
    try:
        exec(compile(src, path, 'exec'), g)
    finally:
        sys.argv = argv
        sys.excepthook = hook
    return g['nt'], g['scene'], g['camera']
