"""The C-ABI shared library loads, exports every symbol include/ntracer_b200.h declares, validates its
arguments like the reference, and FAILS LOUDLY without a GPU (no compute calls are made here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceScene
from tests import fixtures as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ntracer_b200.h')).read()
    return sorted(set(re.findall(r'NTR_API\s+[\w\s\*]+?\b(ntr_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_capi.EXPORTED_SYMBOLS)
    assert lib.ntr_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.Node) == 16
    assert C.sizeof(_capi.Channel) == 20
    assert C.sizeof(_capi.ImageFormat) == 16 + 16 * 20 + 4
    assert C.sizeof(_capi.Counters) == 72


def test_image_format_validation_follows_reference():
    # ImageFormat.__new__ / Channel.__new__ errors, reference src/render.cpp:142-153,202-205,275-280
    with pytest.raises(ValueError):
        _capi.make_image_format(0, 10, _capi.RGB8)
    with pytest.raises(ValueError):
        _capi.make_image_format(10, 10, [(32, 1, 0, 0)])
    with pytest.raises(ValueError):
        _capi.make_image_format(10, 10, [(0, 1, 0, 0)])
    with pytest.raises(ValueError):
        _capi.make_image_format(10, 10, [(8, 1, 0, 0, 0, True)])
    with pytest.raises(ValueError):
        _capi.make_image_format(10, 10, _capi.RGB8, pitch=29)
    with pytest.raises(ValueError):
        _capi.make_image_format(10, 10, [(31, 1, 0, 0)] * 5)
    f = _capi.make_image_format(10, 10, [(5, 1, 0, 0), (6, 0, 1, 0), (5, 0, 0, 1)])
    assert f.bytes_per_pixel == 2 and f.pitch == 20


def test_scene_validation_errors_without_touching_a_device():
    lib = _capi.load()
    sc, g = fx.load('kdtree_kat')
    bad = dict(sc)
    bad['leaf_refs'] = sc['leaf_refs'].copy()
    bad['leaf_refs'][0] = 12345            # simplex index out of range
    desc, keep = _capi.make_desc(bad)
    h = C.c_void_p()
    assert lib.ntr_scene_create(C.byref(desc), -1, C.byref(h)) == _capi.NTR_ERR_VALUE
    assert b'out of range' in lib.ntr_last_error()
    bad = dict(sc)
    bad['dim'] = np.int64(2)
    with pytest.raises(ValueError):         # the arrays no longer fit the record sizes: caught before the library sees them
        _capi.make_desc(bad)
    desc, keep = _capi.make_desc(sc)
    desc.dim = 2
    assert lib.ntr_scene_create(C.byref(desc), -1, C.byref(h)) == _capi.NTR_ERR_VALUE
    # counts without arrays: the C ABI answers NTR_ERR_VALUE, it never dereferences NULL
    for field in ('simplex', 'simplex_mat', 'leaf_refs', 'materials', 'nodes'):
        desc, keep = _capi.make_desc(sc)
        setattr(desc, field, None)
        assert lib.ntr_scene_create(C.byref(desc), -1, C.byref(h)) == _capi.NTR_ERR_VALUE, field
        assert b'NULL' in lib.ntr_last_error(), field
    for missing in ('simplex_mat', 'boundary'):
        bad = {k: v for k, v in sc.items() if k != missing}
        with pytest.raises(ValueError):
            _capi.make_desc(bad)
    # the host-side builders report their own errors
    out = np.zeros(4 * 13, np.float32)
    assert lib.ntr_simplex_from_points(2, 4, None, out.ctypes.data_as(C.c_void_p)) == _capi.NTR_ERR_VALUE
    assert lib.ntr_last_error()


@pytest.mark.skipif(_capi.load().ntr_device_count() > 0, reason='a B200 is present')
def test_no_gpu_means_loud_failure_not_fallback():
    sc, g = fx.load('box4')
    with pytest.raises(_capi.BackendError):
        DeviceScene(sc)
