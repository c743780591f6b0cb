#!/usr/bin/env python3
"""Quick device-time probe for A/B runs of kernel variants (not the benchmark: no oracle counts, no CPU baseline).

  [NTR_B200_LIB=variants/libntr_x.so] python tools/quick.py c4 [--frames 5] [--world 8] [--check]

Prints one JSON line: median / min device ms of a frame of the bench config (ntr_render_device into device memory,
CUDA events of the library), the ray counters, and with --world N the time of rank 0's interleaved share of an
N-GPU frame.  --check compares the float image with the default library's (a second DeviceScene through
NTR_B200_LIB unset is not possible in one process, so the image checksum is printed instead and compared by the
caller)."""
import argparse
import hashlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config')
    ap.add_argument('--frames', type=int, default=5)
    ap.add_argument('--world', type=int, default=1)
    ap.add_argument('--rank', type=int, default=0, help='which of the --world shares to render')
    ap.add_argument('--size', default=None, help='WxH override')
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--chains', default=None, help='device list of a group, e.g. 0,0 = two chains on GPU 0 (ntr_group_*)')
    ap.add_argument('--param', action='append', default=[], help='index=value: overwrite an entry of the scene params (1 = shadows, 3 = max depth)')
    ap.add_argument('--scene-npz', default=None, help='a flat scene dict saved with numpy.savez to render instead of the config\'s fixture (same frame size)')
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from ntracer_b200 import _capi
    from ntracer_b200.backend import DeviceScene
    fixture, w, h, desc = bench.CONFIGS[args.config]
    if args.size:
        w, h = [int(v) for v in args.size.split('x')]
    sc, g = bench.load_fixture(fixture)
    if args.scene_npz:
        z = np.load(args.scene_npz)
        sc = {k: z[k] for k in z.files}
    if args.param:
        sc = dict(sc, params=np.array(sc['params'], dtype=np.float64))
        for kv in args.param:
            k, v = kv.split('=')
            sc['params'][int(k)] = float(v)
    if args.chains:
        from ntracer_b200.backend import DeviceGroup
        grp = DeviceGroup(sc, devices=[int(d) for d in args.chains.split(',')])
        fmt = _capi.make_image_format(w, h, _capi.RGB8)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        ms = []
        for i in range(args.frames + 3):
            flush.zero_()
            torch.cuda.synchronize()
            ptr = grp.render_device(fmt)
            if i >= 3:
                ms.append(grp.last_kernel_ms())
        import ctypes
        host = np.zeros(fmt.pitch * h, np.uint8)
        grp.render(fmt, host)
        print(json.dumps({'config': args.config, 'chains': args.chains, 'w': w, 'h': h, 'ms_median': statistics.median(ms), 'ms_min': min(ms),
                          'ms': [round(v, 3) for v in ms], 'counters': grp.counters(), 'frame_md5': hashlib.md5(host.tobytes()).hexdigest()}), flush=True)
        return
    ds = DeviceScene(sc, 0)
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    rows = ((h + 31) // 32 + args.world - 1) // args.world * 32
    buf = torch.zeros(rows * fmt.pitch if args.world > 1 else fmt.pitch * h, dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ms = []
    for i in range(args.frames + 3):
        flush.zero_()
        torch.cuda.synchronize()
        ds.render_device(fmt, buf.data_ptr(), buf.numel(), st.cuda_stream, args.rank, args.world, args.world > 1)
        st.synchronize()
        if i >= 3:
            ms.append(ds.last_kernel_ms())
    out = {'config': args.config, 'lib': os.environ.get('NTR_B200_LIB', 'default'), 'w': w, 'h': h, 'world': args.world, 'rank': args.rank,
           'ms_median': statistics.median(ms), 'ms_min': min(ms), 'ms': [round(v, 3) for v in ms],
           'counters': ds.counters(), 'frame_md5': hashlib.md5(buf.cpu().numpy().tobytes()).hexdigest()}
    if args.check:
        fl = ds.render_float(min(w, 480), min(h, 270))
        out['float_md5_480x270'] = hashlib.md5(np.round(fl * 255).astype(np.uint8).tobytes()).hexdigest()
        out['counters_480x270'] = ds.counters()
    print(json.dumps(out), flush=True)
    ds.close()


if __name__ == '__main__':
    main()
