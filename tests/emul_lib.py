"""Test-side loader of tests/host_emul/libhostemul.so: the product's per-ray device code
(ntracer_b200/csrc/trace_core.cuh) compiled for the host so that its logic can be checked against the
oracle without a GPU.  Test infrastructure only; the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

from ntracer_b200 import _capi

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_emul')
_SO = os.path.join(_DIR, 'libhostemul.so')
_CSRC = os.path.join(os.path.dirname(_DIR), '..', 'ntracer_b200', 'csrc')
_lib = None


def build():
    gxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([gxx, '-std=c++17', '-O2', '-fPIC', '-shared', '-fvisibility=hidden',
                    '-I/usr/local/cuda/include', '-Wno-unknown-pragmas', '-o', _SO,
                    os.path.join(_DIR, 'emul.cpp')], check=True)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_DIR, 'emul.cpp')] + [os.path.join(_CSRC, f) for f in
                                                    ('trace_core.cuh', 'device_types.h', 'arena_pack.h')]
        if not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def render(sc, w, h, cam=None, generic=False, want_ids=False):
    d, keep = _capi.make_desc(sc)
    if cam is None:
        cam = (sc['cam_origin'], sc['cam_axes'])
    o = np.ascontiguousarray(cam[0], dtype=np.float32)
    a = np.ascontiguousarray(cam[1], dtype=np.float32)
    rgb = np.zeros((h, w, 3), dtype=np.float32)
    ids = np.zeros((h, w), dtype=np.int32)
    dist = np.zeros((h, w), dtype=np.float32)
    cnt = np.zeros(9, dtype=np.uint64)
    lib().emul_render(C.byref(d), _p(o), _p(a), w, h, int(generic), _p(rgb), _p(ids), _p(dist), _p(cnt))
    names = [n for n, _ in _capi.Counters._fields_]
    counters = {n: int(v) for n, v in zip(names, cnt)}
    if want_ids:
        return rgb, ids, dist, counters
    return rgb, counters


def trace_rays(sc, origins, dirs, t_near=-3.4028234663852886e38, t_far=3.4028234663852886e38, skip_ref=None,
               skip_lane=None, generic=False):
    d, keep = _capi.make_desc(sc)
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = origins.shape[0]
    ids = np.zeros(n, dtype=np.int32)
    dist = np.zeros(n, dtype=np.float32)
    nt = np.zeros(n, dtype=np.int32)
    sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
    sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
    lib().emul_trace_rays(C.byref(d), C.c_uint32(n), _p(origins), _p(dirs), C.c_float(t_near), C.c_float(t_far),
                          _p(sr), _p(sl), int(generic), _p(ids), _p(dist), _p(nt))
    return ids, dist, nt


def trace_rays_hits(sc, origins, dirs, t_near=-3.4028234663852886e38, t_far=3.4028234663852886e38, skip_ref=None,
                    skip_lane=None, max_hits=16, generic=False):
    d, keep = _capi.make_desc(sc)
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = origins.shape[0]
    ids = np.zeros(n, dtype=np.int32)
    dist = np.zeros(n, dtype=np.float32)
    nt = np.zeros(n, dtype=np.int32)
    hid = np.full((n, max_hits), -1, dtype=np.int32)
    hdist = np.zeros((n, max_hits), dtype=np.float32)
    sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
    sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
    lib().emul_trace_rays_hits(C.byref(d), C.c_uint32(n), _p(origins), _p(dirs), C.c_float(t_near), C.c_float(t_far),
                               _p(sr), _p(sl), int(generic), _p(ids), _p(dist), _p(nt), int(max_hits), _p(hid), _p(hdist))
    return ids, dist, nt, hid, hdist


def occludes_rays(sc, origins, dirs, distance=None, skip_ref=None, skip_lane=None, generic=False):
    d, keep = _capi.make_desc(sc)
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = origins.shape[0]
    occ = np.zeros(n, dtype=np.int32)
    nt = np.zeros(n, dtype=np.int32)
    dd = None if distance is None else np.ascontiguousarray(distance, dtype=np.float32)
    sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
    sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
    lib().emul_occludes_rays(C.byref(d), C.c_uint32(n), _p(origins), _p(dirs), _p(dd), _p(sr), _p(sl), int(generic),
                             _p(occ), _p(nt))
    return occ, nt


def pack(fmt, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    dst = np.zeros(fmt.pitch * fmt.height, dtype=np.uint8)
    lib().emul_pack(C.byref(fmt), _p(rgb), _p(dst))
    return dst
