"""ntracer_b200 -- B200 (sm_100a) backend for NTracer's per-pixel render loop, behind the reference's Python API.

The names of the reference's `ntracer/__init__.py` are importable from here: Color, Material, Channel, ImageFormat,
CallbackRenderer, BlockingRenderer, NTracer, CUBE, SPHERE.  Rendering always goes through the CUDA library
(ntracer_b200/libntracer_b200.so, C ABI in include/ntracer_b200.h); there is no CPU path."""
from .render import Color, Material, Channel, ImageFormat, CallbackRenderer, BlockingRenderer, StreamRenderer, LockedError  # noqa: F401
from .wrapper import NTracer, CUBE, SPHERE  # noqa: F401
from .backend import DeviceScene, device_count  # noqa: F401
