"""The Python mirror (ntracer_b200.tracern / wrapper) against the compiled reference (oracle/_ref) on random input:
vector / matrix algebra, cross products, camera moves, Triangle.from_points / to_points, prototype bounds.
Skipped where oracle/_ref is not built (it travels with the repository to the GPU box)."""
import math
import os
import random
import sys

import numpy as np
import pytest

import ntracer_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ref_bridge as rb  # noqa: E402

pytestmark = pytest.mark.skipif(not rb.have_reference(), reason='oracle/_ref not built')


def arr(v):
    return np.array(list(v), np.float64)


def marr(m, d):
    return np.array([list(m[i]) for i in range(d)], np.float64)


def close(a, b, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, float(np.abs(b).max()))


@pytest.mark.parametrize('dim', [3, 4, 7])
def test_vector_matrix_camera_algebra(dim):
    rb.load_reference()
    import ntracer as R
    rnd = random.Random(100 + dim)
    rn, mn = R.NTracer(dim), M.NTracer(dim)
    rv = lambda: [rnd.uniform(-3, 3) for _ in range(dim)]
    for trial in range(25):
        a, b = rv(), rv()
        ra, rbv, ma, mb = rn.Vector(a), rn.Vector(b), mn.Vector(a), mn.Vector(b)
        k = rnd.uniform(-2, 2)
        assert close(arr(ma + mb), arr(ra + rbv)) and close(arr(ma - mb), arr(ra - rbv))
        assert close(arr(ma * k), arr(ra * k)) and close(arr(-ma), arr(-ra))
        assert close(arr(ma.unit()), arr(ra.unit())) and close(ma.square(), ra.square()) and close(ma.absolute(), ra.absolute())
        assert close(mn.dot(ma, mb), rn.dot(ra, rbv))
        vs = [rv() for _ in range(dim - 1)]
        assert close(arr(mn.cross([mn.Vector(v) for v in vs])), arr(rn.cross([rn.Vector(v) for v in vs])), 1e-4)
        rows = [tuple(rv()) for _ in range(dim)]
        rows2 = [tuple(rv()) for _ in range(dim)]
        rm, mm, rm2, mm2 = rn.Matrix(rows), mn.Matrix(rows), rn.Matrix(rows2), mn.Matrix(rows2)
        assert close(marr(mm * mm2, dim), marr(rm * rm2, dim), 1e-4)
        assert close(arr(mm * ma), arr(rm * ra), 1e-4)
        assert close(marr(mm.transpose(), dim), marr(rm.transpose(), dim))
        assert close(mm.determinant(), rm.determinant(), 1e-3)
        if abs(rm.determinant()) > 0.5:
            assert close(marr(mm.inverse(), dim), marr(rm.inverse(), dim), 1e-3)
        th = rnd.uniform(-3, 3)
        ua, ub = ra.unit(), rbv.unit()
        assert close(marr(mn.Matrix.rotation(ma.unit(), mb.unit(), th), dim), marr(rn.Matrix.rotation(ua, ub, th), dim), 1e-4)
        assert close(marr(mn.Matrix.scale(k), dim), marr(rn.Matrix.scale(k), dim))
        assert close(marr(mn.Matrix.scale(ma), dim), marr(rn.Matrix.scale(ra), dim))
    # camera: translate / transform / normalize, as scripts/polytope.py drives it
    rc, mc = rn.Camera(), mn.Camera()
    for trial in range(10):
        off = rv()
        rc.translate(rn.Vector(off)); mc.translate(mn.Vector(off))
        i, j = rnd.sample(range(dim), 2)
        th = rnd.uniform(-1, 1)
        rc.transform(rn.Matrix.rotation(rn.Vector.axis(i), rn.Vector.axis(j), th))
        mc.transform(mn.Matrix.rotation(mn.Vector.axis(i), mn.Vector.axis(j), th))
        rc.normalize(); mc.normalize()
        assert close(arr(mc.origin), arr(rc.origin), 1e-4)
        for k in range(dim):
            assert close(arr(mc.axes[k]), arr(rc.axes[k]), 1e-4)
    d = rn.screen_coord_to_ray(rc, 17, 5, 64, 48, 0.8)
    assert close(arr(mn.screen_coord_to_ray(mc, 17, 5, 64, 48, 0.8)), arr(d), 1e-4)


@pytest.mark.parametrize('dim', [3, 5, 6])
def test_triangles_and_prototype_bounds(dim):
    rb.load_reference()
    import ntracer as R
    rnd = random.Random(200 + dim)
    rn, mn = R.NTracer(dim), M.NTracer(dim)
    rmat, mmat = R.Material((1, 1, 1)), M.Material((1, 1, 1))
    for trial in range(20):
        pts = [tuple(rnd.uniform(-10, 10) for _ in range(dim)) for _ in range(dim)]
        rt, mt = rn.Triangle.from_points([rn.Vector(p) for p in pts], rmat), mn.Triangle.from_points([mn.Vector(p) for p in pts], mmat)
        scale = float(np.abs(arr(rt.face_normal)).max())
        assert close(arr(mt.p1), arr(rt.p1)) and close(arr(mt.face_normal) / scale, arr(rt.face_normal) / scale, 1e-3)
        assert close(mt.d / scale, rt.d / scale, 1e-3)
        for e in range(dim - 1):
            assert close(arr(mt.edge_normals[e]), arr(rt.edge_normals[e]), 2e-3)
        for p, q in zip(mt.to_points(), rt.to_points()):
            assert close(arr(p), arr(q), 1e-3)
        rp, mp = rn.TrianglePrototype(pts, rmat), mn.TrianglePrototype(pts, mmat)
        assert close(arr(mp.boundary.start), arr(rp.boundary.start)) and close(arr(mp.boundary.end), arr(rp.boundary.end))
        for j in range(dim):
            assert close(arr(mp.point_data[j].point), arr(rp.point_data[j].point))
            assert close(arr(mp.point_data[j].edge_normal), arr(rp.point_data[j].edge_normal), 2e-3)
        pos = [rnd.uniform(-3, 3) for _ in range(dim)]
        ori = [tuple(rnd.uniform(-1.5, 1.5) for _ in range(dim)) for _ in range(dim)]
        for tr, tm in ((R.CUBE, M.CUBE), (R.SPHERE, M.SPHERE)):
            rs = rn.SolidPrototype(tr, rn.Vector(pos), rn.Matrix(ori), rmat)
            ms = mn.SolidPrototype(tm, mn.Vector(pos), mn.Matrix(ori), mmat)
            assert close(arr(ms.boundary.start), arr(rs.boundary.start), 1e-3), (tr, trial)
            assert close(arr(ms.boundary.end), arr(rs.boundary.end), 1e-3)
            assert close(marr(ms.inv_orientation, dim), marr(rs.inv_orientation, dim), 1e-3)


class _EmulatedDevice:
    """Test double for backend.DeviceScene: the same flat scene, rays traced by the host-emulated device code
    (tests/emul_lib.py).  The product always talks to the CUDA library; this only lets the Python half of
    Primitive.intersects (twins, skip lanes, hit frames of solids) be checked against the reference without a GPU."""
    def __init__(self, sc, device=-1):
        self.sc = sc

    def trace_rays(self, origins, dirs, t_near, t_far, skip_ref=None, skip_lane=None):
        from tests import emul_lib as el
        return el.trace_rays(self.sc, origins, dirs, t_near, t_far, skip_ref, skip_lane)

    def close(self):
        pass


@pytest.mark.parametrize('device', ['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
@pytest.mark.parametrize('dim', [3, 4, 6])
def test_single_primitive_ray_tests(dim, device, monkeypatch):
    """Triangle.intersects / TriangleBatch.intersects / Solid.intersects (src/ntracer_body.hpp:1002-1056) against the
    reference's; once with the host-emulated device code behind the mirror (no GPU), once through the CUDA library."""
    rb.load_reference()
    import ntracer as R
    from ntracer_b200 import tracern
    if device == 'emulated':
        monkeypatch.setattr(tracern, 'DeviceScene', _EmulatedDevice)
    rnd = random.Random(300 + dim)
    rn, mn = R.NTracer(dim), M.NTracer(dim)
    rmat, mmat = R.Material((1, 1, 1), 0.5), M.Material((1, 1, 1), 0.5)           # transparent: still reported
    hits = misses = 0

    def rays(target, n):
        for _ in range(n):
            o = [rnd.uniform(-4, 4) for _ in range(dim)]
            aim = [t + rnd.uniform(-0.6, 0.6) for t in target]
            yield o, [a - b for a, b in zip(aim, o)]

    def same(mh, rh):
        nonlocal hits, misses
        assert (mh is None) == (rh is None)
        if rh is None:
            misses += 1
            return
        hits += 1
        assert close(mh.dist, rh.dist, 1e-4) and close(arr(mh.origin), arr(rh.origin), 1e-3)
        assert close(arr(mh.normal), arr(rh.normal), 1e-3)
        assert mh.batch_index == rh.batch_index

    for trial in range(6):
        pts = [tuple(rnd.uniform(-2, 2) for _ in range(dim)) for _ in range(dim)]
        rt, mt = rn.Triangle.from_points([rn.Vector(p) for p in pts], rmat), mn.Triangle.from_points([mn.Vector(p) for p in pts], mmat)
        centre = [sum(p[i] for p in pts) / dim for i in range(dim)]
        for o, d in rays(centre, 12):
            mh, rh = mt.intersects(mn.Vector(o), mn.Vector(d)), rt.intersects(rn.Vector(o), rn.Vector(d))
            same(mh, rh)
            if mh is not None:
                assert mh.primitive is mt
        # a batch: 4 simplexes around the same centre; skipping the winner exposes the next lane
        ptss = [[tuple(c + rnd.uniform(-1.5, 1.5) for c in centre) for _ in range(dim)] for _ in range(rn.BATCH_SIZE)]
        rbt = rn.TriangleBatch([rn.Triangle.from_points([rn.Vector(p) for p in ps], rmat) for ps in ptss])
        mbt = mn.TriangleBatch([mn.Triangle.from_points([mn.Vector(p) for p in ps], mmat) for ps in ptss])
        for o, d in rays(centre, 12):
            mh, rh = mbt.intersects(mn.Vector(o), mn.Vector(d)), rbt.intersects(rn.Vector(o), rn.Vector(d))
            same(mh, rh)
            if rh is not None:
                same(mbt.intersects(mn.Vector(o), mn.Vector(d), rh.batch_index), rbt.intersects(rn.Vector(o), rn.Vector(d), rh.batch_index))
        pos = [rnd.uniform(-1, 1) for _ in range(dim)]
        ori = [tuple((1.0 if i == j else 0.0) + rnd.uniform(-0.4, 0.4) for j in range(dim)) for i in range(dim)]
        for tr, tm in ((R.CUBE, M.CUBE), (R.SPHERE, M.SPHERE)):
            rs, ms = rn.Solid(tr, rn.Vector(pos), rn.Matrix(ori), rmat), mn.Solid(tm, mn.Vector(pos), mn.Matrix(ori), mmat)
            # the solid sits where inv_orientation*x - position is small (the reference's frame)
            world = list(arr(rn.Matrix(ori) * rn.Vector(pos)))
            for o, d in rays(world, 12):
                same(ms.intersects(mn.Vector(o), mn.Vector(d)), rs.intersects(rn.Vector(o), rn.Vector(d)))
    assert hits > 40 and misses > 10
