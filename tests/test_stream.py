"""The interactive loop: camera path of polytope.py's RotatingCamera and the two-frames-in-flight
begin/end interface.  CPU part: the camera path against the reference's own Camera arithmetic (when oracle/_ref is
built) and its invariants.  GPU part: every streamed frame is bit-identical to the blocking render of the same camera,
and agrees with the oracle."""
import math
import os
import sys

import numpy as np
import pytest

from ntracer_b200 import _capi, stream
from tests import fixtures as fx
from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rotation_cameras_invariants():
    sc, g = fx.load('cell120')
    frames = 12
    cams = stream.rotation_cameras(sc['cam_origin'], sc['cam_axes'], frames)
    assert len(cams) == frames
    assert np.array_equal(cams[0][0], np.asarray(sc['cam_origin'], np.float32))
    assert np.array_equal(cams[0][1], np.asarray(sc['cam_axes'], np.float32))
    dist = float(np.dot(sc['cam_origin'], sc['cam_axes'][2]))
    for o, a in cams[1:]:
        assert np.abs(a @ a.T - np.eye(4)).max() < 1e-5          # orthonormal after Camera.normalize
        assert np.abs(o - a[2] * dist).max() < 1e-5              # back on its orbit, looking at the centre
    # a full turn: one more step from the last camera comes back to the first forward axis
    step = math.acos(max(-1.0, min(1.0, float(np.dot(cams[0][1][2], cams[1][1][2])))))
    assert abs(step - 2 * math.pi / frames) < 1e-3


def test_rotation_cameras_match_the_reference_camera():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_bridge as rb
    if not rb.have_reference():
        pytest.skip('oracle/_ref not built')
    rb.load_reference()
    import ntracer
    for name in ('cell120', 'solids6'):
        sc, g = fx.load(name)
        d = int(sc['dim'])
        frames = 10
        cams = stream.rotation_cameras(sc['cam_origin'], sc['cam_axes'], frames)
        nt = ntracer.NTracer(d)
        cam = nt.Camera()
        cam.origin = nt.Vector(*[float(x) for x in sc['cam_origin']])
        for i in range(d):
            cam.axes[i] = nt.Vector(*[float(x) for x in sc['cam_axes'][i]])
        dist = float(np.dot(sc['cam_origin'], sc['cam_axes'][2]))
        h, incr = 1 / math.sqrt(d - 1), 2 * math.pi / frames
        for k in range(1, frames):                               # scripts/polytope.py:545-555, verbatim semantics
            a2 = cam.axes[0] * h + cam.axes[1] * h
            for i in range(d - 3):
                a2 += cam.axes[i + 3] * h
            cam.transform(nt.Matrix.rotation(cam.axes[2], a2, incr))
            cam.normalize()
            cam.origin = cam.axes[2] * dist
            o = np.array(list(cam.origin), np.float32)
            a = np.array([list(cam.axes[i]) for i in range(d)], np.float32)
            scale = max(1.0, float(np.abs(o).max()))
            assert np.abs(o - cams[k][0]).max() <= 2e-5 * scale
            assert np.abs(a - cams[k][1]).max() <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize('name,variant', [('cell120', None), ('cell120', 'refl'), ('box4', None)])
def test_streamed_frames_equal_blocking_frames_and_the_oracle(name, variant):
    from ntracer_b200.backend import DeviceScene
    sc, g = fx.load(name)
    if variant:
        sc = fx.variant(sc, g, variant)
    w, h = 160, 90
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    cams = stream.rotation_cameras(sc['cam_origin'], sc['cam_axes'], 40)[:9]
    got = {}
    with DeviceScene(sc) as ds:
        bufs = [bytearray(fmt.pitch * h), np.zeros(fmt.pitch * h, np.uint8)]          # pageable: staged path
        n = stream.render_sequence(ds, fmt, cams, bufs, sink=lambda k, b: got.__setitem__(k, np.frombuffer(b, np.uint8).copy()))
        assert n == len(cams) and sorted(got) == list(range(len(cams)))
        for k, (o, a) in enumerate(cams):
            ds.set_camera(o, a)
            assert np.array_equal(ds.render(fmt), got[k]), 'frame %d differs from the blocking render' % k
        for k in (0, 4, 8):
            ref = ol.render_packed(sc, fmt, cam=cams[k])
            d = np.abs(ref.reshape(h, w, 3).astype(np.int32) - got[k].reshape(h, w, 3).astype(np.int32)).max(axis=2)
            assert np.mean(d > 1) <= 0.002


@pytest.mark.gpu
def test_stream_pinned_destination_ticket_rules_and_abort():
    import torch
    from ntracer_b200.backend import DeviceScene
    sc, g = fx.load('cell120')
    w, h = 256, 144
    fmt = _capi.make_image_format(w, h, _capi.RGB8, pitch=w * 3 + 16)
    with DeviceScene(sc) as ds:
        ref = ds.render(fmt, np.full(fmt.pitch * h, 0xAB, np.uint8))
        pinned = [torch.full((fmt.pitch * h,), 0xAB, dtype=torch.uint8).pin_memory() for _ in range(3)]
        t0 = ds.render_begin(fmt, pinned[0].numpy())
        t1 = ds.render_begin(fmt, pinned[1].numpy())
        with pytest.raises(RuntimeError):                       # a third frame: "already running"
            ds.render_begin(fmt, pinned[2].numpy())
        with pytest.raises(ValueError):                         # out of order
            ds.render_end(t1)
        ds.render_end(t0)
        t2 = ds.render_begin(fmt, pinned[2].numpy())
        ds.render_end(t1)
        ds.render_end(t2)
        with pytest.raises(ValueError):
            ds.render_end(t2)                                   # already ended
        for p in pinned:
            assert np.array_equal(p.numpy(), ref)               # direct copy leaves the pitch padding alone too
        with pytest.raises(ValueError):
            ds.render_begin(fmt, bytearray(10))
        # abort hits the frames that are open
        t3 = ds.render_begin(fmt, pinned[0].numpy())
        ds.abort()
        with pytest.raises(_capi.AbortedError):
            ds.render_end(t3)
        t4 = ds.render_begin(fmt, pinned[0].numpy())            # the next frame starts clean
        ds.render_end(t4)
        assert np.array_equal(pinned[0].numpy(), ref)
        # ... also in the steady state of the pipelined loop, where a frame is always open: one abort
        # word per frame -- the abort hits a and b, c is begun afterwards and must come out whole
        for p in pinned:
            p.fill_(0xAB)
        ta = ds.render_begin(fmt, pinned[0].numpy())
        tb = ds.render_begin(fmt, pinned[1].numpy())
        ds.abort()
        with pytest.raises(_capi.AbortedError):
            ds.render_end(ta)
        tc = ds.render_begin(fmt, pinned[2].numpy())            # b is still open
        with pytest.raises(_capi.AbortedError):
            ds.render_end(tb)
        ds.render_end(tc)
        assert np.array_equal(pinned[2].numpy(), ref)
        assert np.array_equal(ds.render(fmt, np.full(fmt.pitch * h, 0xAB, np.uint8)), ref)     # a synchronous call does not care either


def _rotation_golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'rotation.npz'))


@pytest.mark.parametrize('name', ['cell120', 'solids6'])
def test_oracle_and_camera_path_match_reference_frames_of_the_rotation(name):
    """tests/golden/rotation.npz: frames rendered by the reference itself along its RotatingCamera path
    (make_fixtures.py: make_rotation)."""
    g = _rotation_golden()
    sc, _ = fx.load(name)
    frames = int(g['frames'])
    w, h = [int(v) for v in g['size']]
    cams = stream.rotation_cameras(sc['cam_origin'], sc['cam_axes'], frames)
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    for j, k in enumerate(int(v) for v in g['steps']):
        ro, ra = g[name + '_origin'][j], g[name + '_axes'][j]
        assert np.abs(cams[k][0] - ro).max() <= 3e-5 * max(1.0, float(np.abs(ro).max()))
        assert np.abs(cams[k][1] - ra).max() <= 3e-5
        gold = g[name + '_packed'][j].reshape(h, w, 3).astype(np.int32)
        mine = ol.render_packed(sc, fmt, cam=(ro, ra)).reshape(h, w, 3).astype(np.int32)     # the reference's own camera
        assert np.mean(np.abs(mine - gold).max(axis=2) > 1) <= 0.001, (name, k)
        assert len(np.unique(gold.reshape(-1, 3), axis=0)) > 20                                # a real picture


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['cell120', 'solids6'])
def test_streamed_frames_match_reference_frames_of_the_rotation(name):
    from ntracer_b200.backend import DeviceScene
    g = _rotation_golden()
    sc, _ = fx.load(name)
    w, h = [int(v) for v in g['size']]
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    cams = [(g[name + '_origin'][j], g[name + '_axes'][j]) for j in range(len(g['steps']))]
    got = {}
    with DeviceScene(sc) as ds:
        bufs = [np.zeros(fmt.pitch * h, np.uint8), np.zeros(fmt.pitch * h, np.uint8)]
        stream.render_sequence(ds, fmt, cams, bufs, sink=lambda k, b: got.__setitem__(k, b.copy()))
    for j in range(len(cams)):
        gold = g[name + '_packed'][j].reshape(h, w, 3).astype(np.int32)
        d = np.abs(got[j].reshape(h, w, 3).astype(np.int32) - gold).max(axis=2)
        assert np.mean(d > 1) <= 0.001, (name, j)
