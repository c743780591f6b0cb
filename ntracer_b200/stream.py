"""The interactive loop next to the hot path: a camera path rendered frame after frame, the way
scripts/polytope.py does with CallbackRenderer + RotatingCamera (reference scripts/polytope.py:505-557), but with two
frames in flight (ntr_render_begin / ntr_render_end): frame k+1 is traced while frame k is copied to the host."""
import math

import numpy as np

_F = np.float32


def rotation_cameras(cam_origin, cam_axes, frames):
    """The camera path of `polytope.py --benchmark` (RotatingCamera, scripts/polytope.py:522-557): every frame the
    camera is rotated by 2*pi/frames in the plane (forward, h*(right + up + axes[3..])), re-orthonormalised
    (camera.hpp:25-36) and put back at its distance along the new forward axis.  Frame 0 is the given camera.
    -> list of (origin [D], axes [D, D]) float32, arithmetic in float32 like the reference's Camera."""
    from .tracern import Camera, Matrix, Vector
    axes = np.ascontiguousarray(cam_axes, dtype=_F)
    origin = np.ascontiguousarray(cam_origin, dtype=_F)
    d = axes.shape[0]
    cam = Camera(d)
    cam.origin = Vector._wrap(origin.copy())
    for i in range(d):
        cam.axes[i] = Vector._wrap(axes[i].copy())
    cam_distance = float(np.dot(origin, axes[2]))
    incr = 2 * math.pi / frames
    h = 1 / math.sqrt(d - 1)
    out = [(origin.copy(), axes.copy())]
    for _ in range(1, frames):
        a2 = cam.axes[0] * h + cam.axes[1] * h
        for i in range(d - 3):
            a2 = a2 + cam.axes[i + 3] * h
        cam.transform(Matrix.rotation(cam.axes[2], a2, incr))
        cam.normalize()
        cam.origin = cam.axes[2] * cam_distance
        out.append((np.asarray(cam.origin._v, dtype=_F).copy(), np.stack([np.asarray(cam.axes[i]._v, dtype=_F) for i in range(d)])))
    return out


def render_sequence(dev, fmt, cameras, buffers, sink=None):
    """Render one frame per camera into `buffers` (two writable host buffers, used alternately; pinned ones are written
    by the copy engine directly).  `sink(k, buffer)` is called when frame k is complete and before its buffer is reused.
    dev: backend.DeviceScene.  Returns the number of frames rendered."""
    if len(buffers) < 2:
        raise ValueError('two destination buffers are needed to keep two frames in flight')
    pending = None
    n = 0
    for k, (origin, axes) in enumerate(cameras):
        dev.set_camera(origin, axes)
        ticket = dev.render_begin(fmt, buffers[k % 2])
        if pending is not None:
            dev.render_end(pending[0])
            if sink is not None:
                sink(pending[1], buffers[pending[1] % 2])
        pending = (ticket, k)
        n += 1
    if pending is not None:
        dev.render_end(pending[0])
        if sink is not None:
            sink(pending[1], buffers[pending[1] % 2])
    return n
