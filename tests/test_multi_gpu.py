"""Several GPUs behind the C ABI: ntr_group_* (one process, N devices, peer stores into one frame
on the first device) and ntr_frame_* (one process per GPU, the frame shared through CUDA IPC).  The N-device frame must
equal the 1-device frame: byte for byte for single-pass frames, within 1 LSB where float atomics accumulate bounces."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceGroup, DeviceScene, SharedFrame, device_count
from tests import fixtures as fx
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_group_of_one_device_is_the_plain_renderer():
    sc, g = fx.load('cell120')
    w, h = 200, 150                         # ragged last tile row and ragged columns
    fmt = _capi.make_image_format(w, h, _capi.RGB8, pitch=w * 3 + 8)
    with DeviceScene(sc) as ds:
        dest = np.full(fmt.pitch * h, 0xAB, np.uint8)
        ds.render(fmt, dest)
    with DeviceGroup(sc, 1) as grp:
        assert grp.size == 1
        mine = np.full(fmt.pitch * h, 0xAB, np.uint8)
        grp.render(fmt, mine)
        assert np.array_equal(mine, dest)                       # pixels equal, pitch padding untouched (0xAB)
        assert grp.counters()['primary_rays'] == w * h and grp.last_kernel_ms() > 0 and grp.launch_count() >= 1
        with pytest.raises(ValueError):
            grp.render(fmt, bytearray(10))
    for v in ('refl', 'refl_transp'):                           # wavefront passes through the group path
        s2 = fx.variant(sc, g, v)
        with DeviceScene(s2) as ds:
            a = ds.render(fmt).astype(np.int32)
        with DeviceGroup(s2, 1) as grp:
            b = grp.render(fmt).astype(np.int32)
        assert np.abs(a - b).max() <= 1


@pytest.mark.skipif(device_count() < 2, reason='needs at least two B200s')
@pytest.mark.parametrize('n', [2, 3, 4, 8])
def test_group_frame_equals_single_device_frame(n):
    if device_count() < n:
        pytest.skip('needs %d GPUs' % n)
    sc, g = fx.load('cell120')
    w, h = 400, 300
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    with DeviceScene(sc) as ds:
        one = ds.render(fmt)
    with DeviceGroup(sc, n) as grp:
        assert grp.size == n
        many = grp.render(fmt)
        assert np.array_equal(many, one)                        # single pass: byte for byte
        cnt = grp.counters()
        assert cnt['primary_rays'] == w * h
        grp.set_camera(sc['cam_origin'] * 1.1, sc['cam_axes'])  # every device follows the camera
        moved = grp.render(fmt)
    s1 = dict(sc, cam_origin=sc['cam_origin'] * 1.1)
    with DeviceScene(s1) as ds:
        assert np.array_equal(ds.render(fmt), moved)
    s2 = fx.variant(sc, g, 'refl_transp')                       # bounce passes: accumulated with float atomics
    with DeviceScene(s2) as ds:
        one = ds.render(fmt).astype(np.int32)
        fl = ds.render_float(w, h)
    with DeviceGroup(s2, n) as grp:
        many = grp.render(fmt).astype(np.int32)
        assert np.abs(many - one).max() <= 1
        assert grp.counters()['reflection_rays'] > 0
    assert np.abs(many - ol.pack(fmt, fl).astype(np.int32)).max() <= 1


CHILD = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceScene, SharedFrame
from tests import fixtures as fx
sc, g = fx.load('cell120')
if %(variant)r:
    sc = fx.variant(sc, g, %(variant)r)
fmt = _capi.make_image_format(%(w)d, %(h)d, _capi.RGB8)
frame = SharedFrame(%(device)d, handle=bytes.fromhex(%(handle)r))
with DeviceScene(sc, %(device)d) as ds:
    ds.render_device(fmt, frame.ptr, fmt.pitch * fmt.height, 0, 1, 2, False)      # rank 1 of 2, rows at their frame position
    import torch
    torch.cuda.synchronize(%(device)d)
frame.close()
print('child done')
'''


@pytest.mark.parametrize('variant', ['', 'refl'])
def test_two_processes_store_into_one_shared_frame(variant):
    """The torchrun arrangement of bench.py without torchrun: this process is rank 0 and owns the frame, a child process
    is rank 1 (on the second GPU when there is one, else on the same GPU), maps the frame through its IPC handle and
    stores its tile rows into it.  Rows of the other rank must be left untouched by each (sentinel), also when a pack
    kernel follows wavefront passes."""
    import torch
    sc, g = fx.load('cell120')
    if variant:
        sc = fx.variant(sc, g, variant)
    w, h = 200, 150
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    nbytes = fmt.pitch * h
    with DeviceScene(sc, 0) as ds:
        full = ds.render(fmt).reshape(h, fmt.pitch).astype(np.int32)
        frame = SharedFrame(0, nbytes)
        try:
            frame.fill(0x5A, nbytes)
            torch.cuda.synchronize(0)
            ds.render_device(fmt, frame.ptr, nbytes, 0, 0, 2, False)                # rank 0 of 2
            torch.cuda.synchronize(0)
            half = np.zeros(nbytes, np.uint8)
            frame.download(fmt, half)
            torch.cuda.synchronize(0)
            rows = half.reshape(h, fmt.pitch)
            for ty in range((h + 31) // 32):
                blk = rows[ty * 32:(ty + 1) * 32].astype(np.int32)
                if ty % 2 == 0:
                    assert np.abs(blk - full[ty * 32:(ty + 1) * 32]).max() <= (1 if variant else 0)
                else:
                    assert np.all(blk == 0x5A)                                      # not this rank's rows: untouched
            child_dev = 1 if device_count() >= 2 else 0
            code = CHILD % {'root': ROOT, 'variant': variant, 'w': w, 'h': h, 'device': child_dev, 'handle': frame.export().hex()}
            out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, cwd=ROOT)
            assert 'child done' in out.stdout, (out.stdout[-1000:], out.stderr[-2000:])
            both = np.zeros(nbytes, np.uint8)
            frame.download(fmt, both)
            torch.cuda.synchronize(0)
            assert np.abs(both.reshape(h, fmt.pitch).astype(np.int32) - full).max() <= (1 if variant else 0)
        finally:
            frame.close()


@pytest.mark.skipif(device_count() < 2, reason='needs at least two B200s')
def test_blocking_renderer_threads_is_the_gpu_count():
    """BlockingRenderer(threads=N) through the mirror of the reference's API: N GPUs trace the frame (render.py: _gpus_for)."""
    from ntracer_b200 import render as R
    from ntracer_b200.wrapper import NTracer
    nt = NTracer(4)
    scene = nt.BoxScene()
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -5))
    scene.set_camera(cam)
    fmt = R.ImageFormat(320, 240, [R.Channel(8, 1, 0, 0), R.Channel(8, 0, 1, 0), R.Channel(8, 0, 0, 1)])
    one, two = bytearray(320 * 240 * 3), bytearray(320 * 240 * 3)
    assert R.BlockingRenderer().render(one, fmt, scene)
    r2 = R.BlockingRenderer(2)
    assert r2._gpus == 2 and r2.render(two, fmt, scene)
    assert one == two
