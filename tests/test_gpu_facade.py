"""End-to-end through the reference-compatible Python API on a B200 (mirrors how scripts/hypercube.py and
scripts/polytope.py drive the reference)."""
import threading

import numpy as np
import pytest

from ntracer_b200 import BlockingRenderer, CallbackRenderer, Channel, ImageFormat, Material, NTracer
from ntracer_b200 import _capi, tracern
from tests import fixtures as fx
from tests import oracle_lib as ol
from tests.test_facade import kat_scene

pytestmark = pytest.mark.gpu
RGB = [Channel(8, 1, 0, 0), Channel(8, 0, 1, 0), Channel(8, 0, 0, 1)]


def test_boxscene_like_hypercube_py():
    nt = NTracer(4)
    scene = nt.BoxScene()
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -5))
    scene.set_camera(cam)
    fmt = ImageFormat(640, 480, RGB)
    buf = bytearray(640 * 480 * 3)
    assert BlockingRenderer().render(buf, fmt, scene) is True
    sc, g = fx.load('box4')
    d = np.abs(np.frombuffer(buf, np.uint8).astype(np.int32) - g['packed'].astype(np.int32))
    assert d.max() <= 1 and np.count_nonzero(d) <= 12
    c = scene.calculate_color(100, 200, 640, 480)
    assert list(c) == pytest.approx([0, 0.27875945, 0.27875945], abs=3e-7)
    assert scene.locked == 0


def test_reference_test_kdtree_through_the_api():
    nt = NTracer(3)
    scene, prims = kat_scene(nt)
    hits = scene.root.intersects((4.917067527770996, 2.508934497833252, -4.304379940032959),
                                 (-0.7135500907897949, -0.1356230527162552, 0.6873518228530884))
    assert len(hits) == 1
    assert prims.index(hits[0].primitive) == 4
    assert hits[0].batch_index == -1
    occ, _ = scene.root.occludes((4.917067527770996, 2.508934497833252, -4.304379940032959),
                                 (-0.7135500907897949, -0.1356230527162552, 0.6873518228530884))
    assert occ is True


def test_built_scene_renders_like_the_oracle_and_callback_renderer():
    nt = NTracer(4)
    rng = np.random.RandomState(7)
    mats = [Material((1, 0.5, 0.5)), Material((0.3, 0.8, 1.0), 1, 0.4)]
    protos = []
    for i in range(150):
        c = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.05, 0.05, 1)])
        protos.append(nt.TrianglePrototype([nt.Vector(*(c + rng.uniform(-0.3, 0.3, 4))) for _ in range(4)], mats[i % 2]))
    scene = nt.build_composite_scene(protos)
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -4))
    scene.set_camera(cam)
    scene.add_light(nt.GlobalLight(nt.Vector(0.2, -1, 0.3, 0), (0.7, 0.7, 0.7)))
    w, h = 96, 64
    fmt = ImageFormat(w, h, RGB)
    buf = bytearray(w * h * 3)
    done = threading.Event()
    r = CallbackRenderer()
    r.begin_render(buf, fmt, scene, lambda rr: done.set())
    assert done.wait(60)
    r.abort_render()
    flat = tracern._Flattener(4)
    root = flat.walk(scene.root)
    sc = flat.scene_dict(root, scene.boundary, scene)
    sc['cam_origin'], sc['cam_axes'] = scene._cam._origin, scene._cam._axes
    o = ol.render_packed(sc, _capi.make_image_format(w, h, _capi.RGB8)).reshape(h, w, 3).astype(np.int32)
    a = np.frombuffer(buf, np.uint8).reshape(h, w, 3).astype(np.int32)
    assert np.mean(np.abs(a - o).max(axis=2) > 1) <= 0.002
    # Material objects are live like in the reference: changing one changes the next frame
    before = bytes(buf)
    mats[0].color = (0.1, 0.9, 0.1)
    assert BlockingRenderer().render(buf, fmt, scene)
    assert bytes(buf) != before
