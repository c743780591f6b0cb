"""DeviceScene: thin object wrapper over the C ABI (include/ntracer_b200.h).

One DeviceScene = one scene arena resident on one B200.  The reference-compatible classes in
ntracer_b200.tracern / ntracer_b200.render build flat scene dicts and drive this class; tests and
bench.py use it directly.  Every method goes through libntracer_b200.so -- there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _capi

FLT_MAX = 3.4028234663852886e38


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class DeviceScene:
    def __init__(self, scene, device=-1):
        """scene: flat scene dict (DESIGN.md 'scene file'; ntracer_b200.scene_io.load_scene)."""
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.dim = int(scene['dim'])
        self.kind = int(scene['kind'])
        self._scene = scene
        self._open = {}
        desc, keep = _capi.make_desc(scene)
        _capi.check(self._lib.ntr_scene_create(C.byref(desc), device, C.byref(self._h)))
        if 'cam_origin' in scene and 'cam_axes' in scene:
            self.set_camera(scene['cam_origin'], scene['cam_axes'])

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.ntr_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state ----
    def set_camera(self, origin, axes):
        o = np.ascontiguousarray(origin, dtype=np.float32).reshape(self.dim)
        a = np.ascontiguousarray(axes, dtype=np.float32).reshape(self.dim, self.dim)
        _capi.check(self._lib.ntr_scene_set_camera(self._h, _p(o), _p(a)))

    def set_params(self, scene):
        """Re-send fov / shadows / camera_light / max_reflect_depth / background / ambient / lights."""
        desc, keep = _capi.make_desc(scene)
        _capi.check(self._lib.ntr_scene_set_params(self._h, C.byref(desc)))
        self._scene = scene

    def set_instrumented(self, on):
        _capi.check(self._lib.ntr_set_instrumented(self._h, int(bool(on))))

    # ---- the hot path ----
    def render(self, fmt, dest=None):
        """BlockingRenderer.render: packs the frame into `dest` (any writable buffer) or a new uint8 array."""
        need = fmt.pitch * fmt.height
        if dest is None:
            out = np.zeros(need, dtype=np.uint8)
            buf = out
        else:
            out = dest
            buf = np.frombuffer(dest, dtype=np.uint8)
            if buf.size < need:
                raise ValueError('the buffer is too small for an image with the given dimensions')
        _capi.check(self._lib.ntr_render(self._h, C.byref(fmt), _p(buf), buf.size))
        return out

    def render_begin(self, fmt, dest):
        """Enqueue one frame with the current camera and return a ticket (the interactive loop: frame k+1 is traced
        while frame k is copied back; at most two frames open).  `dest` must stay alive until render_end(ticket)."""
        buf = np.frombuffer(dest, dtype=np.uint8)
        if buf.size < fmt.pitch * fmt.height:
            raise ValueError('the buffer is too small for an image with the given dimensions')
        ticket = C.c_uint64(0)
        _capi.check(self._lib.ntr_render_begin(self._h, C.byref(fmt), _p(buf), buf.size, C.byref(ticket)))
        self._open[ticket.value] = buf          # keeps the exported buffer alive while the copy engine writes it
        return ticket.value

    def render_end(self, ticket):
        """Wait for the frame of `ticket` and finish the copy into its destination."""
        try:
            _capi.check(self._lib.ntr_render_end(self._h, C.c_uint64(ticket)))
        finally:
            self._open.pop(ticket, None)

    def render_device(self, fmt, dev_ptr, nbytes, stream=0, tile_row_first=0, tile_row_step=1, compact=False):
        """Asynchronous render into device memory (a torch tensor's data_ptr()) on a CUDA stream handle."""
        _capi.check(self._lib.ntr_render_device(self._h, C.byref(fmt), C.c_void_p(dev_ptr), nbytes,
                                                C.c_void_p(stream), tile_row_first, tile_row_step, int(compact)))

    def render_float(self, width, height):
        out = np.zeros((height, width, 3), dtype=np.float32)
        _capi.check(self._lib.ntr_render_float(self._h, width, height, _p(out)))
        return out

    def calculate_color(self, x, y, width, height):
        out = (C.c_float * 3)()
        _capi.check(self._lib.ntr_calculate_color(self._h, x, y, width, height, out))
        return np.array(list(out), dtype=np.float32)

    def primary_hit_ids(self, width, height):
        ids = np.zeros((height, width), dtype=np.int32)
        dist = np.zeros((height, width), dtype=np.float32)
        _capi.check(self._lib.ntr_primary_hit_ids(self._h, width, height, _p(ids), _p(dist)))
        return ids, dist

    def trace_rays(self, origins, dirs, t_near=-FLT_MAX, t_far=FLT_MAX, skip_ref=None, skip_lane=None):
        origins = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, self.dim)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, self.dim)
        n = origins.shape[0]
        ids = np.zeros(n, dtype=np.int32)
        dist = np.zeros(n, dtype=np.float32)
        nt = np.zeros(n, dtype=np.int32)
        sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
        sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
        _capi.check(self._lib.ntr_trace_rays(self._h, n, _p(origins), _p(dirs), t_near, t_far, _p(sr), _p(sl),
                                             _p(ids), _p(dist), _p(nt)))
        return ids, dist, nt

    def trace_rays_hits(self, origins, dirs, t_near=-FLT_MAX, t_far=FLT_MAX, skip_ref=None, skip_lane=None, max_hits=16):
        """trace_rays plus the surviving transparent hits of every ray: (ids, dist, n_transparent, hit_ids [n, max_hits]
        (-1 = unused), hit_dists [n, max_hits]) -- what KDNode.intersects returns ahead of the opaque hit."""
        origins = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, self.dim)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, self.dim)
        n = origins.shape[0]
        ids = np.zeros(n, dtype=np.int32)
        dist = np.zeros(n, dtype=np.float32)
        nt = np.zeros(n, dtype=np.int32)
        hid = np.full((n, max_hits), -1, dtype=np.int32)
        hdist = np.zeros((n, max_hits), dtype=np.float32)
        sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
        sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
        _capi.check(self._lib.ntr_trace_rays_hits(self._h, n, _p(origins), _p(dirs), t_near, t_far, _p(sr), _p(sl),
                                                  _p(ids), _p(dist), _p(nt), int(max_hits), _p(hid), _p(hdist)))
        return ids, dist, nt, hid, hdist

    def occludes_rays(self, origins, dirs, distance=None, skip_ref=None, skip_lane=None):
        origins = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, self.dim)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, self.dim)
        n = origins.shape[0]
        occ = np.zeros(n, dtype=np.int32)
        nt = np.zeros(n, dtype=np.int32)
        dd = None if distance is None else np.ascontiguousarray(distance, dtype=np.float32)
        sr = None if skip_ref is None else np.ascontiguousarray(skip_ref, dtype=np.uint32)
        sl = None if skip_lane is None else np.ascontiguousarray(skip_lane, dtype=np.int32)
        _capi.check(self._lib.ntr_occludes_rays(self._h, n, _p(origins), _p(dirs), _p(dd), _p(sr), _p(sl),
                                                _p(occ), _p(nt)))
        return occ, nt

    # ---- control / introspection ----
    def abort(self):
        _capi.check(self._lib.ntr_abort(self._h))

    def counters(self):
        c = _capi.Counters()
        _capi.check(self._lib.ntr_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def last_kernel_ms(self):
        ms = C.c_float()
        _capi.check(self._lib.ntr_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        return int(self._lib.ntr_launch_count(self._h))


class DeviceGroup:
    """One scene replicated on several B200s of one box (ntr_group_*): a frame is split by interleaved 32-pixel tile rows
    and every device stores its rows straight into one frame buffer on the first device over NVLink."""
    def __init__(self, scene, n=0, devices=None):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.dim = int(scene['dim'])
        desc, keep = _capi.make_desc(scene)
        devs = None if devices is None else (C.c_int * len(devices))(*[int(d) for d in devices])
        _capi.check(self._lib.ntr_group_create(C.byref(desc), len(devices) if devices is not None else int(n), devs, C.byref(self._h)))
        self.size = int(self._lib.ntr_group_size(self._h))
        if 'cam_origin' in scene and 'cam_axes' in scene:
            self.set_camera(scene['cam_origin'], scene['cam_axes'])

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.ntr_group_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_camera(self, origin, axes):
        o = np.ascontiguousarray(origin, dtype=np.float32).reshape(self.dim)
        a = np.ascontiguousarray(axes, dtype=np.float32).reshape(self.dim, self.dim)
        _capi.check(self._lib.ntr_group_set_camera(self._h, _p(o), _p(a)))

    def set_params(self, scene):
        desc, keep = _capi.make_desc(scene)
        _capi.check(self._lib.ntr_group_set_params(self._h, C.byref(desc)))

    def render(self, fmt, dest=None):
        need = fmt.pitch * fmt.height
        out = np.zeros(need, dtype=np.uint8) if dest is None else dest
        buf = out if dest is None else np.frombuffer(dest, dtype=np.uint8)
        if buf.size < need:
            raise ValueError('the buffer is too small for an image with the given dimensions')
        _capi.check(self._lib.ntr_group_render(self._h, C.byref(fmt), _p(buf), buf.size))
        return out

    def render_device(self, fmt):
        """-> address of the frame on the first device (valid until the next render)"""
        ptr = C.c_void_p()
        _capi.check(self._lib.ntr_group_render_device(self._h, C.byref(fmt), C.byref(ptr)))
        return ptr.value

    def abort(self):
        _capi.check(self._lib.ntr_group_abort(self._h))

    def counters(self):
        c = _capi.Counters()
        _capi.check(self._lib.ntr_group_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def last_kernel_ms(self):
        ms = C.c_float()
        _capi.check(self._lib.ntr_group_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        return int(self._lib.ntr_group_launch_count(self._h))


class SharedFrame:
    """A frame buffer in the memory of one GPU that kernels of OTHER processes store into (one process per GPU):
    rank 0 allocates and exports it, the others import the handle (CUDA IPC)."""
    def __init__(self, device, nbytes=0, handle=None):
        self._lib = _capi.load()
        self.device, self.owner = int(device), handle is None
        ptr = C.c_void_p()
        if self.owner:
            _capi.check(self._lib.ntr_frame_alloc(self.device, int(nbytes), C.byref(ptr)))
        else:
            h = (C.c_ubyte * 64).from_buffer_copy(bytes(handle))
            _capi.check(self._lib.ntr_frame_import(self.device, h, C.byref(ptr)))
        self.ptr = ptr.value

    def export(self):
        h = (C.c_ubyte * 64)()
        _capi.check(self._lib.ntr_frame_export(C.c_void_p(self.ptr), h))
        return bytes(h)

    def fill(self, value, nbytes, stream=0):
        _capi.check(self._lib.ntr_frame_fill(self.device, C.c_void_p(self.ptr), int(value), int(nbytes), C.c_void_p(stream)))

    def download(self, fmt, dest, stream=0):
        buf = np.frombuffer(dest, dtype=np.uint8)
        _capi.check(self._lib.ntr_frame_download(self.device, C.c_void_p(self.ptr), C.byref(fmt), _p(buf), buf.size, C.c_void_p(stream)))

    def close(self):
        if self.ptr:
            (self._lib.ntr_frame_free if self.owner else self._lib.ntr_frame_release)(self.device, C.c_void_p(self.ptr))
            self.ptr = None


def device_count():
    return int(_capi.load().ntr_device_count())


def measure_fp32_peak(device=-1):
    out = C.c_float()
    _capi.check(_capi.load().ntr_measure_fp32_peak(device, C.byref(out)))
    return float(out.value)
