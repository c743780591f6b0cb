import sys, time, json
import numpy as np
from ntracer_b200 import bulk, _capi
from ntracer_b200.backend import DeviceScene
for n, depth, w, h in ((20000, 0, 960, 540), (200000, 0, 960, 540), (200000, 16, 960, 540), (1000000, 14, 960, 540), (1000000, 17, 960, 540), (1000000, 20, 960, 540)):
    pts = bulk.soup(10, n)
    t = time.time(); sc = bulk.simplex_scene(pts, max_depth=depth); tb = time.time() - t
    sc['cam_origin'] = np.array([0, 0, -3] + [0] * 7, np.float32)
    leaf = (sc['nodes'][:, 0] & 0x80000000) != 0
    t = time.time(); ds = DeviceScene(sc); tu = time.time() - t
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    img = ds.render(fmt); img = ds.render(fmt)
    ms = ds.last_kernel_ms()
    ids, dist = ds.primary_hit_ids(w, h)
    print(json.dumps({'n': n, 'max_depth': depth, 'build_s': round(tb, 1), 'upload_s': round(tu, 2), 'nodes': int(len(sc['nodes'])), 'refs': int(len(sc['leaf_refs'])),
                      'max_leaf': int(sc['nodes'][leaf][:, 2].max()), 'kernel_ms_%dx%d' % (w, h): round(ms, 2), 'Mpix_s': round(w * h / ms / 1e3, 2), 'hit_frac': round(float((ids >= 0).mean()), 3)}), flush=True)
    ds.close()
