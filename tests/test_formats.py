"""Data formats either side of the path: pickling of the mirror's objects (the reference's test_pickle,
lib/ntracer/tests/test.py:365-382) and the Wavefront .obj reader (lib/ntracer/wavefront_obj.py)."""
import os
import pickle
import random
import sys

import numpy as np
import pytest

from ntracer_b200 import Color, Material, NTracer, wavefront_obj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def roundtrip(x):
    return pickle.loads(pickle.dumps(x))


def test_pickle_roundtrips_like_the_reference_test():
    rnd = random.Random(4)
    mat = Material((1, 1, 1))
    assert roundtrip(mat) == mat
    assert roundtrip(Color(0.2, 0.1, 1)) == Color(0.2, 0.1, 1)
    for d in (3, 5, 12):
        nt = NTracer(d)
        rv = lambda lo=-1000, hi=1000: nt.Vector([rnd.uniform(lo, hi) for _ in range(d)])
        v = rv()
        assert roundtrip(v) == v and type(roundtrip(v)) is nt.base.Vector
        a = nt.AABB(rv(-100, 50), rv(51, 200))
        b = roundtrip(a)
        assert b.start == a.start and b.end == a.end
        t = nt.Triangle(rv(), rv(), [rv() for _ in range(d - 1)], mat)
        u = roundtrip(t)
        assert (u.p1, u.face_normal, u.d, u.edge_normals, u.material) == (t.p1, t.face_normal, t.d, t.edge_normals, t.material)
        m = nt.Matrix.identity()
        assert roundtrip(m) == m
        c = nt.Camera()
        c.translate(nt.Vector.axis(2, -5))
        assert roundtrip(c).origin == c.origin


def test_pickled_scene_keeps_tree_and_state():
    nt = NTracer(3)
    mat = Material((1, 0.5, 0.5), 0.5, 0.25)
    protos = [nt.TrianglePrototype([(0, 0, z), (1, 0, z), (0, 1, z)], mat) for z in (0.0, 0.5, 1.0)]
    scene = nt.build_composite_scene(protos)
    scene.set_fov(1.1)
    scene.set_shadows(True)
    scene.add_light(nt.PointLight(nt.Vector(1, 2, 3), Color(4, 5, 6)))
    scene.locked = 1                       # as if a render were in flight in this process
    copy = roundtrip(scene)
    scene.locked = 0
    assert copy.locked == 0 and copy.fov == scene.fov and copy.shadows is True
    assert copy.boundary.start == scene.boundary.start and copy.boundary.end == scene.boundary.end
    assert len(copy.point_lights) == 1 and copy.point_lights[0].position == nt.Vector(1, 2, 3)
    flat = lambda n: [(type(n).__name__, len(n))] if isinstance(n, nt.KDLeaf) else \
        [(type(n).__name__, n.axis, n.split)] + (flat(n.left) if n.left else []) + (flat(n.right) if n.right else [])
    assert flat(copy.root) == flat(scene.root)


OBJ = """# a unit square made of one quad, one triangle with texture/normal indices, one relative face
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0 1.0
vt 0 0
vn 0 0 1
f 1 2 3 4
f 1/1/1 2/1/1 3//1
v 0 0 1
f -1 -2 -3
"""


def test_obj_reader(tmp_path):
    path = tmp_path / 'square.obj'
    path.write_text(OBJ)
    tris = wavefront_obj.load_obj(str(path))
    assert len(tris) == 4                                     # quad -> 2, triangle, relative triangle
    nt = NTracer(3)
    pts = [[tuple(round(c, 6) for c in pd.point) for pd in t.point_data] for t in tris]
    assert pts[0] == [(0, 0, 0), (1, 0, 0), (1, 1, 0)] and pts[1] == [(0, 0, 0), (1, 1, 0), (0, 1, 0)]
    assert pts[2] == [(0, 0, 0), (1, 0, 0), (1, 1, 0)]
    assert pts[3] == [(0, 0, 1), (0, 1, 0), (1, 1, 0)]
    assert all(isinstance(t, nt.TrianglePrototype) for t in tris)
    scene = nt.build_composite_scene(tris)                    # what a script does next
    assert scene.boundary.end[0] >= 1
    with pytest.raises(ValueError):
        wavefront_obj.load_obj(str(path), NTracer(4))
    for bad in ('v 1 2\n', 'v a b c\n', 'v 0 0 0\nf 1 2 3\n', 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 x\n', 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n'):
        p = tmp_path / 'bad.obj'
        p.write_text(bad)
        with pytest.raises(wavefront_obj.FileFormatError):
            wavefront_obj.load_obj(str(p))


def test_obj_reader_matches_the_reference(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_bridge as rb
    if not rb.have_reference():
        pytest.skip('oracle/_ref not built')
    rb.load_reference()
    from ntracer import wavefront_obj as ref_obj
    rnd = random.Random(9)
    lines = ['v %f %f %f' % (rnd.uniform(-2, 2), rnd.uniform(-2, 2), rnd.uniform(-2, 2)) for _ in range(30)]
    for _ in range(25):
        k = rnd.choice((3, 3, 4, 5))
        idx = rnd.sample(range(1, 31), k)
        lines.append('f ' + ' '.join(('%d/%d' % (i, i)) if rnd.random() < 0.3 else str(i if rnd.random() < 0.7 else i - 31) for i in idx))
    path = tmp_path / 'mesh.obj'
    path.write_text('\n'.join(lines) + '\n')
    mine = wavefront_obj.load_obj(str(path))
    theirs = ref_obj.load_obj(str(path))
    assert len(mine) == len(theirs) > 25
    for a, b in zip(mine, theirs):
        pa = np.array([list(pd.point) for pd in a.point_data], np.float32)
        pb = np.array([list(pd.point) for pd in b.point_data], np.float32)
        assert np.array_equal(pa, pb)
        assert np.allclose(list(a.face_normal), list(b.face_normal), rtol=1e-5, atol=1e-6)
