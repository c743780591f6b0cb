"""The façade tests of tests/test_gpu_facade.py (hypercube.py's BoxScene frame, the reference's test_kdtree through the
API, a built scene through BlockingRenderer / CallbackRenderer) once more on a box without a GPU: the CUDA library
behind the mirror is replaced by a test double that answers from the host-emulated device code, so that everything on
the Python side of the C ABI (wrappers, flattening, renderers, locking, callbacks) is exercised in the CPU tier too.
The product itself never falls back to this: without the library or a device it raises (tests/test_capi.py)."""
import ctypes as C

import pytest

from tests import emul_lib as el
from tests import oracle_lib as ol
from tests import test_gpu_facade as gpu_tests
from tests.test_mirror_vs_reference import _EmulatedRenderDevice


class _FakeLibrary:
    def __init__(self, dev):
        self.dev = dev

    def ntr_render(self, handle, fmt_ref, ptr, size):
        fmt = fmt_ref._obj
        packed = ol.pack(fmt, el.render(self.dev.sc, fmt.width, fmt.height)[0])
        addr = ptr.value if hasattr(ptr, 'value') else ptr
        row_bytes = fmt.width * fmt.bytes_per_pixel
        for y in range(fmt.height):                 # pixel bytes only, like the real call
            row = packed[y * fmt.pitch:y * fmt.pitch + row_bytes]
            C.memmove(addr + y * fmt.pitch, row.ctypes.data, row_bytes)
        return 0


class _EmulatedFullDevice(_EmulatedRenderDevice):
    def __init__(self, sc, device=-1):
        super().__init__(sc, device)
        self._lib, self._h = _FakeLibrary(self), C.c_void_p(1)

    def calculate_color(self, x, y, w, h):
        return el.render(self.sc, w, h)[0][y, x]

    def occludes_rays(self, origins, dirs, distance, skip_ref=None, skip_lane=None):
        return el.occludes_rays(self.sc, origins, dirs, distance, skip_ref, skip_lane)

    def counters(self):
        return {}

    def abort(self):
        pass


@pytest.mark.parametrize('name', ['test_boxscene_like_hypercube_py', 'test_reference_test_kdtree_through_the_api',
                                  'test_built_scene_renders_like_the_oracle_and_callback_renderer'])
def test_gpu_facade_test_on_the_emulated_device(name, monkeypatch):
    from ntracer_b200 import tracern
    monkeypatch.setattr(tracern, 'DeviceScene', _EmulatedFullDevice)
    getattr(gpu_tests, name)()
