"""The boundary from the reference's own host language: examples/cpp_host/render_frame.cpp is a C++ program on
include/ntracer_b200.h (the calls INTEGRATION.md section A.3 adds to src/render.cpp).  CPU tier: it compiles against the
header, links against the in-tree library and fails loudly without a device; GPU tier: its frame equals the one the
ctypes path produces."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / 'render_frame')
    gxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([gxx, '-std=c++17', '-Wall', '-Werror', '-I' + os.path.join(ROOT, 'include'),
                    os.path.join(ROOT, 'examples', 'cpp_host', 'render_frame.cpp'), '-L' + os.path.join(ROOT, 'ntracer_b200'),
                    '-lntracer_b200', '-Wl,-rpath,' + os.path.join(ROOT, 'ntracer_b200'), '-o', exe], check=True)
    return exe


def test_cpp_host_builds_links_and_fails_loudly_without_a_device(tmp_path):
    from ntracer_b200 import _capi
    exe = build(tmp_path)
    if _capi.load().ntr_device_count() > 0:
        pytest.skip('a B200 is present: see the gpu test')
    out = subprocess.run([exe, '4', '64', '48', str(tmp_path / 'f.rgb')], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3 and 'no CPU fallback' in out.stderr          # NTR_ERR_NO_DEVICE, the library's own message
    assert not os.path.exists(tmp_path / 'f.rgb')


@pytest.mark.gpu
def test_cpp_host_frame_equals_the_ctypes_frame(tmp_path):
    from ntracer_b200 import _capi
    from ntracer_b200.backend import DeviceScene, device_count
    from tests import fixtures as fx
    exe = build(tmp_path)
    sc, g = fx.load('box4')
    w, h = 640, 480
    with DeviceScene(sc) as ds:
        mine = ds.render(_capi.make_image_format(w, h, _capi.RGB8))
    for gpus in ([1, 2] if device_count() >= 2 else [1]):
        path = str(tmp_path / ('f%d.rgb' % gpus))
        out = subprocess.run([exe, '4', str(w), str(h), path, str(gpus)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert np.array_equal(np.fromfile(path, np.uint8), mine)
        assert 'kernel launch' in out.stdout
