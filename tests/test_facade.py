"""The host-side mirror of the reference's Python API (ntracer_b200.{wrapper,tracern,render}): the reference's
own unit tests that touch the render path, restated (lib/ntracer/tests/test.py), plus flattening checks.
CPU only: nothing here renders (that is tests/test_gpu_facade.py)."""
import math

import numpy as np
import pytest

from ntracer_b200 import BlockingRenderer, Channel, Color, ImageFormat, Material, NTracer, CUBE, SPHERE
from ntracer_b200 import tracern
from tests import fixtures as fx
from tests import oracle_lib as ol


def kat_scene(nt):
    """The scene of the reference's test_kdtree (lib/ntracer/tests/test.py:302-363)."""
    mat = Material((1, 1, 1))
    T = nt.Triangle
    prims = [
        T((-1.1755770444869995, 0.3819499611854553, -1.6180520057678223), (1.7082732915878296, -2.3512351512908936, 1.4531432390213013),
          [(-0.615524172782898, -0.3236003816127777, 0.19999605417251587), (0.49796950817108154, 0.0381958931684494, -0.5235964059829712)], mat),
        T((-1.1755770444869995, 0.3819499611854553, -1.6180520057678223), (1.0557708740234375, -1.4531433582305908, 0.8980922102928162),
          [(-0.8057316541671753, -0.06180214881896973, 0.8471965789794922), (0.19020742177963257, -0.2617982029914856, -0.6472004652023315)], mat),
        T((0.7265498042106628, 0.9999955296516418, 1.6180428266525269), (0, 1.7961481809616089, 0.8980742692947388),
          [(-1.1135050058364868, -0.1618017703294754, 0.32360348105430603), (0.6881839036941528, -0.09999901801347733, 0.19999800622463226)], mat),
        T((0.7265498042106628, 0.9999955296516418, 1.6180428266525269), (0, 2.90622878074646, 1.4531147480010986),
          [(-0.4253210127353668, -0.26180076599121094, 0.5236014127731323), (0.6881839036941528, 0.09999898821115494, -0.1999979317188263)], mat),
        T((1.9021340608596802, 0.618022620677948, -0.3819592595100403), (-1.055770754814148, -1.4531432390213013, 0.8980920910835266),
          [(-0.30776214599609375, -0.42359834909439087, -1.0471925735473633), (0.4979696571826935, -0.038195837289094925, 0.5235962867736816)], mat),
        T((1.9021340608596802, 0.618022620677948, -0.3819592595100403), (-1.7082730531692505, -2.3512353897094727, 1.4531434774398804),
          [(0.19020749628543854, -0.4617941677570343, -0.5235962271690369), (0.19020745158195496, 0.2617981433868408, 0.6472005844116211)], mat)]
    scene = nt.CompositeScene(
        nt.AABB((-1.710653305053711e-05, 0.618022620677948, -0.3819774389266968), (0.7265291213989258, 2.000016689300537, 0.3819882869720459)),
        nt.KDBranch(1, 2.0000057220458984,
                    nt.KDBranch(1, 0.9999955296516418, None,
                                nt.KDLeaf([prims[4], prims[5], prims[2], prims[3], prims[1], prims[0]])),
                    nt.KDLeaf([prims[4], prims[5], prims[1], prims[0]])))
    scene.set_fov(0.8)
    return scene, prims


def test_math():                                    # reference test_math (test.py:120-130)
    nt = NTracer(4)
    ma = nt.Matrix([[10, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 16]])
    mb = nt.Matrix([13, 6, 9, 6, 7, 3, 3, 13, 1, 11, 12, 7, 12, 15, 17, 15])
    mx = ma * mb
    my = nt.Matrix([195, 159, 200, 167, 210, 245, 283, 277, 342, 385, 447, 441, 474, 525, 611, 605])
    assert list(mx.values) == pytest.approx(list(my.values))
    for a, b in zip((mb * mb.inverse()).values, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]):
        assert a == pytest.approx(b, abs=1e-3)
    assert nt.Vector(13, 2, 16, 14).unit()[0] == pytest.approx(0.52, abs=1e-2)


def test_vector_and_wrapper_api():
    nt = NTracer(7)
    assert NTracer(7) is nt                          # cached per dimension (wrapper.py:100-110)
    v = nt.Vector(1, 2, 3, 4, 5, 6, 7)
    assert list(v) == list(memoryview(v))            # reference test_buffer_interface
    c = Color(0.5, 0.1, 0)
    assert list(c) == pytest.approx(list(memoryview(c)))
    assert nt.dot(v, v) == pytest.approx(140)
    assert abs(v) == pytest.approx(math.sqrt(140))
    assert (v * 2 - v) == v
    assert nt.Vector.axis(2, -5) == nt.Vector(0, 0, -5, 0, 0, 0, 0)
    with pytest.raises(TypeError):
        nt.Vector(1, 2, 3)
    with pytest.raises(ValueError):
        NTracer(2)
    for d in (64, 32):                               # above NTR_MAX_DIM: refused up front, no silent fallback
        with pytest.raises(ValueError):
            NTracer(d)


def test_to_from_points_roundtrip():                 # reference test_to_from_points (test.py:399-406), dimension 5
    nt = NTracer(5)
    rng = np.random.RandomState(4)
    pts = []
    for i in range(5):
        pts.append(nt.Vector(*([rng.uniform(-10, 10) for _ in range(i)] + [rng.uniform(1, 10)] + [0] * (4 - i))))
    back = nt.Triangle.from_points(pts, Material((1, 1, 1))).to_points()
    for a, b in zip(pts, back):
        assert list(a) == pytest.approx(list(b), abs=1e-3)


def test_camera_matches_reference_semantics():
    nt = NTracer(4)
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -5))
    assert cam.origin == nt.Vector(0, 0, -5, 0)
    cam.transform(nt.Matrix.rotation(nt.Vector.axis(0), nt.Vector.axis(2), 0.3))
    cam.normalize()
    a = np.stack([cam.axes[i]._v for i in range(4)])
    assert np.allclose(a @ a.T, np.eye(4), atol=1e-6)
    # screen_coord_to_ray golden from SURVEY.md 8(c)
    d = nt.screen_coord_to_ray(nt.Camera(), 100, 200, 640, 480, 0.8)
    assert list(d) == pytest.approx([-0.27875945, 0.05068354, 0.95902264, 0], abs=1e-6)


def test_scene_locking_and_validation():
    nt = NTracer(3)
    scene, prims = kat_scene(nt)
    scene.locked += 1
    from ntracer_b200 import LockedError
    with pytest.raises(LockedError):
        scene.set_fov(1.0)
    with pytest.raises(LockedError):
        scene.add_light(nt.PointLight((0, 0, 0), (1, 1, 1)))
    scene.locked -= 1
    scene.set_fov(1.0)
    with pytest.raises(TypeError):
        nt.KDBranch(0, 1.0)
    with pytest.raises(ValueError):
        Channel(32, 1, 0, 0)
    with pytest.raises(ValueError):
        ImageFormat(10, 10, [Channel(8, 1, 0, 0)], pitch=5)
    f = ImageFormat(10, 10, [Channel(5, 1, 0, 0), Channel(6, 0, 1, 0), Channel(5, 0, 0, 1)])
    assert f.bytes_per_pixel == 2 and f.pitch == 20
    with pytest.raises(ValueError):
        BlockingRenderer().render(bytearray(10), f, scene)          # buffer too small (render.cpp:187-190)
    with pytest.raises(BufferError):
        BlockingRenderer().render(bytes(400), f, scene)             # not writable


def test_flattened_kat_scene_equals_the_exported_reference_scene():
    nt = NTracer(3)
    scene, prims = kat_scene(nt)
    flat = tracern._Flattener(3)
    root = flat.walk(scene.root)
    sc = flat.scene_dict(root, scene.boundary, scene)
    ref, g = fx.load('kdtree_kat')                  # exported from the real reference by make_fixtures.py
    for k in ('nodes', 'leaf_refs', 'simplex_mat', 'boundary'):
        assert np.array_equal(sc[k], ref[k]), k
    assert np.allclose(sc['simplex'], ref['simplex'], rtol=0, atol=1e-6)     # d = -dot(fn,p1) recomputed on the host
    # and the oracle finds primitive 4 on the flattened scene
    ids, dist, nt_ = ol.trace_rays(sc, g['origin'][None], g['direction'][None])
    owner, lane = flat.prim_of_flat_id(int(ids[0]))
    assert owner is prims[4] and lane == -1


def test_builder_tree_gives_reference_hit_ids():
    """build_composite_scene (own builder) on random simplexes: hit ids / colours from the oracle must not depend
    on the tree (compare with a single-leaf tree over the same primitives)."""
    nt = NTracer(4)
    rng = np.random.RandomState(7)
    mat = Material((1, 0.5, 0.5))
    protos = []
    for i in range(150):
        c = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.05, 0.05, 1)])
        protos.append(nt.TrianglePrototype([nt.Vector(*(c + rng.uniform(-0.3, 0.3, 4))) for _ in range(4)], mat))
    scene = nt.build_composite_scene(protos)
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -4))
    scene.set_camera(cam)
    one = nt.CompositeScene(scene.boundary, nt.KDLeaf([p.primitive for p in protos]))
    one.set_camera(cam)
    out = []
    for s in (scene, one):
        flat = tracern._Flattener(4)
        root = flat.walk(s.root)
        sc = flat.scene_dict(root, s.boundary, s)
        sc['cam_origin'], sc['cam_axes'] = s._cam._origin, s._cam._axes
        ids, dist = ol.primary_hit_ids(sc, 64, 48)
        plist = [p.primitive for p in protos]

        def owner(i):                   # the Triangle that was hit: a single primitive, or the lane of a TriangleBatch
            prim, lane = flat.prim_of_flat_id(int(i))
            return plist.index(prim[lane] if lane >= 0 else prim)
        owners = np.array([-1 if i < 0 else owner(i) for i in ids.ravel()])
        out.append((owners, dist.ravel(), ol.render_float(sc, 64, 48)))
    assert np.mean(out[0][0] == out[1][0]) >= 0.999
    assert np.abs(out[0][2] - out[1][2]).max() < 1e-4
    assert (out[0][0] >= 0).mean() > 0.1
    # the built tree holds the triangles as TriangleBatch items (group_primitives, src/tracer.hpp:2395-2427)
    def leaves(n):
        return [n] if isinstance(n, tracern.KDLeaf) else [l for c in (n.left, n.right) if c is not None for l in leaves(c)]
    items = {id(it): it for l in leaves(scene.root) for it in l}
    batches = [it for it in items.values() if isinstance(it, tracern.TriangleBatch)]
    singles = [it for it in items.values() if isinstance(it, tracern.Triangle)]
    assert len(batches) == 150 // 4 and len(singles) == 150 % 4
    assert sorted(plist.index(t) for b in batches for t in b) + sorted(plist.index(t) for t in singles) == sorted(range(150)) or \
        sorted([plist.index(t) for b in batches for t in b] + [plist.index(t) for t in singles]) == list(range(150))


# ---- the builder-side geometry tests of the reference's own test-suite (lib/ntracer/tests/test.py), same vectors -------
def test_aabb_reference_vectors():                   # test.py:132-140
    nt = NTracer(5)
    a = nt.AABB((1, 7, -5, 5, 4), (5, 13, -1, 6, 12))
    assert a.dimension == 5
    assert list(a.end) == [5, 13, -1, 6, 12] and list(a.start) == [1, 7, -5, 5, 4]
    assert list(a.right(2, -3).start) == [1, 7, -3, 5, 4]
    assert list(a.left(0, 2).end) == [2, 13, -1, 6, 12]


def test_aabb_triangle_reference_vectors():          # test.py:142-203
    nt = NTracer(3)
    mat = Material((1, 1, 1))
    box = nt.AABB((-1, -1, -1), (1, 1, 1))
    tri = lambda pts: nt.TrianglePrototype(pts, mat)
    assert not box.intersects(tri([(-2.092357, 0.1627209, 0.9231308), (0.274588, 0.8528936, 2.309217), (-1.212236, 1.855952, 0.3137006)]))
    assert not box.intersects(tri([(2.048058, -3.022543, 1.447644), (1.961913, -0.5438575, -0.1552723), (0.3618142, -1.684767, 0.2162201)]))
    assert not box.intersects(tri([(-4.335572, -1.690142, -1.302721), (0.8976227, 0.5090631, 4.6815), (-0.8176082, 4.334341, -1.763081)]))
    assert box.intersects(tri([(0, 0, 0), (5, 5, 5), (1, 2, 3)]))
    assert nt.AABB((-0.894424974918, -1.0, -0.850639998913), (0.0, -0.447214990854, 0.850639998913)).intersects(
        tri([(0.0, -1.0, 0.0), (0.723599970341, -0.447214990854, 0.525720000267), (-0.276385009289, -0.447214990854, 0.850639998913)]))
    rng = np.random.RandomState(3)
    points = [[tuple(rng.uniform(-1, 1, 3)) for _ in range(3)] for _ in range(nt.BATCH_SIZE)]
    flat = np.array(points, np.float32).reshape(-1, 3)
    tbp = nt.TriangleBatchPrototype(tri(p) for p in points)
    assert np.allclose(list(tbp.boundary.start), flat.min(axis=0)) and np.allclose(list(tbp.boundary.end), flat.max(axis=0))
    assert nt.BATCH_SIZE == 4
    assert box.intersects(nt.TriangleBatchPrototype([
        tri([(5.8737568855285645, 0.0, 0.0), (2.362654209136963, 1.4457907676696777, 0.0), (-7.4159417152404785, -2.368093252182007, 5.305923938751221)]),
        tri([(6.069871425628662, 0.0, 0.0), (8.298105239868164, 1.4387503862380981, 0.0), (-7.501928806304932, 4.3413987159729, 5.4995622634887695)]),
        tri([(5.153589248657227, 0.0, 0.0), (-0.8880055546760559, 3.595335006713867, 0.0), (-0.14510761201381683, 6.0621466636657715, 1.7603594064712524)]),
        tri([(1.9743329286575317, 0.0, 0.0), (-0.6579152345657349, 8.780682563781738, 0.0), (1.0433781147003174, 0.5538825988769531, 4.187061309814453)])]))


def test_aabb_cube_and_sphere_reference_vectors():   # test.py:205-267
    from ntracer_b200 import CUBE, SPHERE
    nt = NTracer(3)
    mat = Material((1, 1, 1))
    box = nt.AABB((-1, -1, -1), (1, 1, 1))
    cube = lambda pos, m: nt.SolidPrototype(CUBE, nt.Vector(*pos), nt.Matrix(*m), mat)
    assert not box.intersects(cube((1.356136, 1.717844, 1.577731),
                                   (-0.01922399, -0.3460019, 0.8615935, -0.03032121, -0.6326356, -0.5065715, 0.03728577, -0.6928598, 0.03227519)))
    assert not box.intersects(cube((1.444041, 1.433598, 1.975453),
                                   (0.3780299, -0.3535482, 0.8556266, -0.7643852, -0.6406123, 0.07301452, 0.5223108, -0.6816301, -0.5124177)))
    assert not box.intersects(cube((-0.31218, -3.436678, 1.473133),
                                   (0.8241131, -0.2224413, 1.540015, -1.461101, -0.7099018, 0.6793453, 0.5350775, -1.595884, -0.516849)))
    assert not box.intersects(cube((0.7697315, -3.758033, 1.847144),
                                   (0.6002195, -1.608681, -0.3900863, -1.461104, -0.7098908, 0.6793506, -0.7779449, 0.0921175, -1.576897)))
    assert box.intersects(cube((0.4581598, -1.56134, 0.5541568),
                               (0.3780299, -0.3535482, 0.8556266, -0.7643852, -0.6406123, 0.07301452, 0.5223108, -0.6816301, -0.5124177)))
    assert not box.intersects(nt.SolidPrototype(SPHERE, nt.Vector(-1.32138, 1.6959, 1.729396), nt.Matrix.identity(), mat))
    assert box.intersects(nt.SolidPrototype(SPHERE, nt.Vector(1.623511, -1.521197, -1.243952), nt.Matrix.identity(), mat))
    with pytest.raises(TypeError):
        box.intersects(nt.Solid(SPHERE, nt.Vector(0, 0, 0), nt.Matrix.identity(), mat))       # primitives need a prototype
    with pytest.raises(TypeError):
        NTracer(4).AABB().intersects(nt.SolidPrototype(SPHERE, nt.Vector(0, 0, 0), nt.Matrix.identity(), mat))


def test_batch_and_buffer_interfaces():              # test.py:269-300
    nt = NTracer(4)
    rng = np.random.RandomState(11)
    lo, hi = (lambda: float(rng.uniform(-1, 1))), (lambda: float(rng.uniform(9, 11)))
    protos = [nt.TrianglePrototype([(lo(), lo(), lo(), lo()), (lo(), hi(), lo(), lo()), (hi(), lo(), lo(), lo()), (lo(), lo(), hi(), lo())],
                                   Material((1, 1, 1.0 / (i + 1)))) for i in range(nt.BATCH_SIZE)]
    bproto = nt.TriangleBatchPrototype(protos)
    for i in range(nt.BATCH_SIZE):
        assert protos[i].face_normal == bproto.face_normal[i]
        for j in range(nt.dimension):
            assert protos[i].point_data[j].point == bproto.point_data[j].point[i]
            assert protos[i].point_data[j].edge_normal == bproto.point_data[j].edge_normal[i]
        assert protos[i].material == bproto.material[i]
    v = NTracer(7).Vector(1, 2, 3, 4, 5, 6, 7)
    assert list(v) == list(memoryview(v))
    c = Color(0.5, 0.1, 0)
    assert list(c) == list(memoryview(c))


def test_aabb_tests_agree_with_the_reference_on_random_input():
    """AABB.intersects / intersects_flat for simplexes, batches, cubes and spheres against the compiled reference."""
    import os, sys, random
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import ref_bridge as rb
    if not rb.have_reference():
        pytest.skip('oracle/_ref not built')
    rb.load_reference()
    import ntracer as R
    import ntracer_b200 as M
    random.seed(5)
    seen = set()
    for dim in (3, 4, 6):
        rn, mn = R.NTracer(dim), M.NTracer(dim)
        rmat, mmat = R.Material((1, 1, 1)), M.Material((1, 1, 1))
        for trial in range(120):
            lo = tuple(random.uniform(-2, 1) for _ in range(dim))
            hi = tuple(l + random.uniform(0.1, 2.5) for l in lo)
            rbx, mbx = rn.AABB(lo, hi), mn.AABB(lo, hi)
            pts = [tuple(random.uniform(-3, 3) for _ in range(dim)) for _ in range(dim)]
            rp, mp = rn.TrianglePrototype(pts, rmat), mn.TrianglePrototype(pts, mmat)
            sk = random.randrange(dim)
            pos = tuple(random.uniform(-3, 3) for _ in range(dim))
            ori = [tuple(random.uniform(-1.5, 1.5) for _ in range(dim)) for _ in range(dim)]
            got = [mbx.intersects(mp), mbx.intersects_flat(mp, sk)]
            want = [rbx.intersects(rp), rbx.intersects_flat(rp, sk)]
            for typ_r, typ_m in ((R.CUBE, M.CUBE), (R.SPHERE, M.SPHERE)):
                want.append(rbx.intersects(rn.SolidPrototype(typ_r, rn.Vector(pos), rn.Matrix(ori), rmat)))
                got.append(mbx.intersects(mn.SolidPrototype(typ_m, mn.Vector(pos), mn.Matrix(ori), mmat)))
            if trial % 4 == 0:
                ptss = [[tuple(random.uniform(-3, 3) for _ in range(dim)) for _ in range(dim)] for _ in range(rn.BATCH_SIZE)]
                rb_, mb_ = (rn.TriangleBatchPrototype([rn.TrianglePrototype(p, rmat) for p in ptss]),
                            mn.TriangleBatchPrototype([mn.TrianglePrototype(p, mmat) for p in ptss]))
                want += [rbx.intersects(rb_), rbx.intersects_flat(rb_, sk)]
                got += [mbx.intersects(mb_), mbx.intersects_flat(mb_, sk)]
            assert got == want, (dim, trial)
            seen.update(want)
    assert seen == {True, False}
