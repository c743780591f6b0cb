"""Test-side loader for oracle/liboracle.so (the C restatement of the reference; test infrastructure).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from ntracer_b200 import _capi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, 'oracle')
_SO = os.path.join(_ORACLE_DIR, 'liboracle.so')
_lib = None


def build():
    subprocess.run(['make', '-C', _ORACLE_DIR, '-s'], check=True)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_ORACLE_DIR, 'ntr_oracle.c')
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _cam(sc, cam=None):
    if cam is None:
        cam = (sc['cam_origin'], sc['cam_axes'])
    o = np.ascontiguousarray(cam[0], dtype=np.float32)
    a = np.ascontiguousarray(cam[1], dtype=np.float32)
    return o, a


def render_float(sc, w, h, cam=None, window=None, with_mask=False, with_counters=False):
    d, keep = _capi.make_desc(sc)
    o, a = _cam(sc, cam)
    rgb = np.zeros((h, w, 3), dtype=np.float32)
    mask = np.zeros((h, w), dtype=np.uint8)
    cnt = _capi.Counters()
    x0, y0, x1, y1 = window if window else (0, 0, w, h)
    lib().oracle_render_float_window(C.byref(d), _p(o), _p(a), w, h, x0, y0, x1, y1, _p(rgb), _p(mask), C.byref(cnt))
    out = [rgb]
    if with_mask:
        out.append(mask)
    if with_counters:
        out.append(cnt.as_dict())
    return out[0] if len(out) == 1 else tuple(out)


def calculate_color(sc, x, y, w, h, cam=None):
    d, keep = _capi.make_desc(sc)
    o, a = _cam(sc, cam)
    out = (C.c_float * 3)()
    lib().oracle_calculate_color(C.byref(d), _p(o), _p(a), x, y, w, h, out)
    return np.array(list(out), dtype=np.float32)


def pack(fmt, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    dst = np.zeros(fmt.pitch * fmt.height, dtype=np.uint8)
    lib().oracle_pack(C.byref(fmt), _p(rgb), _p(dst))
    return dst


def render_packed(sc, fmt, cam=None):
    d, keep = _capi.make_desc(sc)
    o, a = _cam(sc, cam)
    dst = np.zeros(fmt.pitch * fmt.height, dtype=np.uint8)
    lib().oracle_render_packed(C.byref(d), _p(o), _p(a), C.byref(fmt), _p(dst), None)
    return dst


def primary_hit_ids(sc, w, h, cam=None):
    d, keep = _capi.make_desc(sc)
    o, a = _cam(sc, cam)
    ids = np.zeros((h, w), dtype=np.int32)
    dist = np.zeros((h, w), dtype=np.float32)
    lib().oracle_primary_hit_ids(C.byref(d), _p(o), _p(a), w, h, _p(ids), _p(dist))
    return ids, dist


def trace_rays(sc, origins, dirs, t_near=-3.4028234663852886e38, t_far=3.4028234663852886e38, skip_ref=None, skip_lane=None):
    d, keep = _capi.make_desc(sc)
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = origins.shape[0]
    ids = np.zeros(n, dtype=np.int32)
    dist = np.zeros(n, dtype=np.float32)
    nt = np.zeros(n, dtype=np.int32)
    sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
    sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
    lib().oracle_trace_rays(C.byref(d), C.c_uint32(n), _p(origins), _p(dirs), C.c_float(t_near), C.c_float(t_far),
                            _p(sr), _p(sl), _p(ids), _p(dist), _p(nt))
    return ids, dist, nt


def occludes_rays(sc, origins, dirs, distance=None, skip_ref=None, skip_lane=None):
    d, keep = _capi.make_desc(sc)
    origins = np.ascontiguousarray(origins, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32)
    n = origins.shape[0]
    occ = np.zeros(n, dtype=np.int32)
    nt = np.zeros(n, dtype=np.int32)
    dd = None if distance is None else np.ascontiguousarray(distance, dtype=np.float32)
    sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
    sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
    lib().oracle_occludes_rays(C.byref(d), C.c_uint32(n), _p(origins), _p(dirs), _p(dd), _p(sr), _p(sl), _p(occ), _p(nt))
    return occ, nt


def screen_coord_to_ray(dim, cam_axes, x, y, w, h, fov):
    a = np.ascontiguousarray(cam_axes, dtype=np.float32)
    out = np.zeros(dim, dtype=np.float32)
    lib().oracle_screen_coord_to_ray(dim, _p(a), C.c_float(x), C.c_float(y), w, h, C.c_float(fov), _p(out))
    return out


def trace_ray_full(sc, origin, direction, t_near=-3.4028234663852886e38, t_far=3.4028234663852886e38,
                   skip_ref=0xFFFFFFFF, skip_lane=-1, max_hits=64):
    """KDNode.intersects for one ray with full hit records: list of (dist, flat id, origin, normal)."""
    d, keep = _capi.make_desc(sc)
    dim = int(sc['dim'])
    o = np.ascontiguousarray(origin, dtype=np.float32)
    di = np.ascontiguousarray(direction, dtype=np.float32)
    dist = np.zeros(max_hits, np.float32)
    ids = np.zeros(max_hits, np.int32)
    po = np.zeros((max_hits, dim), np.float32)
    pn = np.zeros((max_hits, dim), np.float32)
    n = lib().oracle_trace_ray_full(C.byref(d), _p(o), _p(di), C.c_float(t_near), C.c_float(t_far),
                                    C.c_uint32(skip_ref), skip_lane, max_hits, _p(dist), _p(ids), _p(po), _p(pn))
    return [(float(dist[i]), int(ids[i]), po[i].copy(), pn[i].copy()) for i in range(n)]
