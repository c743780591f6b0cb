"""Per-pixel cost map of a fixture scene from the oracle's counters (test infrastructure: analysis only).
usage: python tools/ray_cost_map.py ggs120[:variant] [width height]
Prints the distribution of simplex tests per pixel (all rays of the pixel: primary, shadow, bounces) -- the quantity
behind the per-pass latency floor discussed in DESIGN.md section 7/8."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import fixtures as fx, oracle_lib as ol    # noqa: E402


def main():
    name, _, var = sys.argv[1].partition(':')
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (160, 90)
    sc, g = fx.load(name)
    if var:
        sc = fx.variant(sc, g, var)
    cost = np.zeros((h, w), np.int64)
    nodes = np.zeros((h, w), np.int64)
    for y in range(h):
        for x in range(w):
            _, cnt = ol.render_float(sc, w, h, window=(x, y, x + 1, y + 1), with_counters=True)
            cost[y, x] = cnt['simplex_tests'] + cnt['solid_tests']
            nodes[y, x] = cnt['node_steps']
    flat = np.sort(cost.ravel())[::-1]
    tot = flat.sum()
    print('%s %dx%d: tests per pixel  mean %.0f  median %.0f  p99 %.0f  max %d' %
          (sys.argv[1], w, h, flat.mean(), np.median(flat), np.percentile(flat, 99), flat[0]))
    for frac in (0.001, 0.01, 0.05, 0.10):
        k = max(1, int(len(flat) * frac))
        print('  the most expensive %.1f %% of pixels hold %.1f %% of all tests (>= %d tests each)' % (100 * frac, 100 * flat[:k].sum() / tot, flat[k - 1]))
    print('  max / mean = %.1f   (a frame split over P warps cannot finish faster than max/(mean*pixels/P) of its ideal time)' % (flat[0] / flat.mean()))
    np.save(os.path.join(ROOT, 'gpurun_out', 'cost_%s_%dx%d.npy' % (name, w, h)), cost)


if __name__ == '__main__':
    main()
