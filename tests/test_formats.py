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


def test_pickle_payloads_are_the_reference_encodings():
    """Known answers taken from the compiled reference (src/render.cpp:1391-1657): every value type reduces to
    render._<type>_unpickle with big-endian IEEE-754 floats."""
    from ntracer_b200 import render as R, tracern as T
    nt = NTracer(3)
    f, a = nt.Vector(1, 2, 3).__reduce__()
    assert f is R._vector_unpickle and a == (3, b'?\x80\x00\x00@\x00\x00\x00@@\x00\x00')
    f, a = Color(0.2, 0.1, 1).__reduce__()
    assert f is R._color_unpickle and a == (b'>L\xcc\xcd=\xcc\xcc\xcd?\x80\x00\x00',)
    mat = Material((1, 0.5, 0.25), 0.5, 0.25)
    f, a = mat.__reduce__()
    assert f is R._material_unpickle and a == (b'?\x80\x00\x00?\x00\x00\x00>\x80\x00\x00?\x80\x00\x00?\x80\x00\x00?\x80\x00\x00?\x00\x00\x00>\x80\x00\x00?\x80\x00\x00A\x00\x00\x00',)
    f, a = nt.AABB(nt.Vector(0, 0, 0), nt.Vector(1, 1, 1)).__reduce__()
    assert f is R._aabb_unpickle and a == (3, b'\x00' * 12 + b'?\x80\x00\x00' * 3)
    f, a = nt.Matrix.identity().__reduce__()
    assert f is R._matrix_unpickle and a[0] == 3 and np.array_equal(np.frombuffer(a[1], '>f4').reshape(3, 3), np.eye(3))
    tri = nt.Triangle.from_points([(0, 0, 0), (1, 0, 0), (0, 1, 0)], mat)
    f, a = tri.__reduce__()
    assert f is R._triangle_unpickle and a[0] == 3 and a[2] is mat
    assert np.array_equal(np.frombuffer(a[1], '>f4').reshape(4, 3), [[0, 0, 0], [0, 0, 1], [-1, -1, 0], [0, 0, 0]][:2] + [[-1, 0, 0], [0, -1, 0]])
    tb = nt.TriangleBatch([nt.Triangle.from_points([(0, 0, z), (1, 0, z), (0, 1, z)], mat) for z in range(4)])
    f, a = tb.__reduce__()
    assert f is R._triangle_batch_unpickle and a[:2] == (4, 3) and a[3:] == (mat,) * 4
    assert np.array_equal(np.frombuffer(a[2], '>f4').reshape(4, 12),                       # [row][coordinate][lane]
                          [[0] * 8 + [0, 1, 2, 3], [0] * 8 + [1] * 4, [-1] * 4 + [0] * 8, [0] * 4 + [-1] * 4 + [0] * 4])
    sol = nt.Solid(T.CUBE, nt.Vector(1, 2, 3), nt.Matrix.identity(), mat)
    f, a = sol.__reduce__()
    assert f is R._solid_unpickle and a[0] == 3 and a[1][:1] == b'\x01' and len(a[1]) == 4 * 12 + 1 and a[2] is mat
    assert np.array_equal(np.frombuffer(a[1][1:], '>f4'), list(np.eye(3).ravel()) + [1, 2, 3])
    # round trips through the encodings, and their error behaviour
    for x in (tb, sol, tri):
        y = roundtrip(x)
        assert type(y) is type(x) and y.__reduce__()[1][:-1 if x is not tb else 3] == x.__reduce__()[1][:-1 if x is not tb else 3]
    assert roundtrip(sol).inv_orientation == sol.inv_orientation and roundtrip(tb)[2].d == tb[2].d
    with pytest.raises(ValueError, match='vector data is malformed'):
        R._vector_unpickle(3, b'\x00' * 8)
    with pytest.raises(TypeError, match='takes exactly 2 arguments'):
        R._vector_unpickle(3)
    with pytest.raises(ValueError, match='color data is malformed'):
        R._color_unpickle(b'abc')
    with pytest.raises(TypeError, match='different batch size'):
        R._triangle_batch_unpickle(8, 3, b'', mat)
    with pytest.raises(ValueError, match='solid data is malformed'):
        R._solid_unpickle(3, b'\x01', mat)
    with pytest.raises(ValueError):
        R._aabb_unpickle(2, b'')


def test_pickles_cross_between_this_package_and_the_reference():
    """A pickle written by the compiled reference loads here (ntracer_b200.compat.loads_reference) and the other way
    round (dumps_for_reference, every protocol): same functions, same payload, only the module name differs."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_bridge as rb
    if not rb.have_reference():
        pytest.skip('oracle/_ref not built')
    rb.load_reference()
    import ntracer as R
    import ntracer_b200 as M
    from ntracer_b200 import compat
    rnd = random.Random(9)
    for d in (3, 5, 12):
        rn, mn = R.NTracer(d), M.NTracer(d)
        vals = lambda: [rnd.uniform(-100, 100) for _ in range(d)]
        rmat, mmat = R.Material((1, 0.5, 0.25), 0.5, 0.25, 2, 9, (0.1, 0.2, 0.3)), M.Material((1, 0.5, 0.25), 0.5, 0.25, 2, 9, (0.1, 0.2, 0.3))
        v = vals()
        rows = [vals() for _ in range(d)]
        pairs = [(rn.Vector(v), mn.Vector(v)), (rn.Matrix(rows), mn.Matrix(rows)), (R.Color(0.2, 0.1, 1), M.Color(0.2, 0.1, 1)), (rmat, mmat),
                 (rn.AABB(rn.Vector(v), rn.Vector([x + 1 for x in v])), mn.AABB(mn.Vector(v), mn.Vector([x + 1 for x in v]))),
                 (rn.Triangle(rn.Vector(v), rn.Vector(rows[0]), [rn.Vector(r) for r in rows[1:]], rmat),
                  mn.Triangle(mn.Vector(v), mn.Vector(rows[0]), [mn.Vector(r) for r in rows[1:]], mmat)),
                 (rn.Solid(R.CUBE, rn.Vector(v), rn.Matrix(rows), rmat), mn.Solid(M.CUBE, mn.Vector(v), mn.Matrix(rows), mmat))]
        rt, mt = pairs[5]
        pairs.append((rn.TriangleBatch([rt] * 4), mn.TriangleBatch([mt] * 4)))
        plain = lambda args: tuple(a for a in args if isinstance(a, (int, bytes)))
        for r, m in pairs:
            assert plain(r.__reduce__()[1]) == plain(m.__reduce__()[1]), type(m).__name__       # the same bytes
            here = compat.loads_reference(pickle.dumps(r))
            assert type(here) is type(m) and plain(here.__reduce__()[1]) == plain(m.__reduce__()[1])
            for proto in (2, 3, 4, 5):
                there = pickle.loads(compat.dumps_for_reference(m, proto))
                assert type(there) is type(r) and plain(there.__reduce__()[1]) == plain(r.__reduce__()[1])
        # hyperspheres: the reference writes them but refuses to read them back (`data[0] != 1 && data[1] != 2`,
        # src/render.cpp:1621, looks at the first byte of the orientation); they load here
        rs, ms = rn.Solid(R.SPHERE, rn.Vector(v), rn.Matrix(rows), rmat), mn.Solid(M.SPHERE, mn.Vector(v), mn.Matrix(rows), mmat)
        assert plain(rs.__reduce__()[1]) == plain(ms.__reduce__()[1])
        assert compat.loads_reference(pickle.dumps(rs)).type == M.SPHERE
        mats = compat.loads_reference(pickle.dumps(pairs[-1][0])).__reduce__()[1][3:]
        assert len(mats) == 4 and all(x == mmat for x in mats)


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
