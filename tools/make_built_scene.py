#!/usr/bin/env python3
"""Rebuilds a polytope fixture's scene on THIS repo's tree: the vertices of every simplex are recovered from its record
(edge_normal_i . (p1 - p_j) = delta_ij, src/tracer.hpp:454-461), grouped into 4-lane batches (ntr_group_items) and handed to
the culled builder (ntr_build_kdtree_culled) -- what nt.build_composite_scene(prototypes) does -- and the flat scene is saved
for `tools/quick.py <config> --scene-npz out.npz`.  CPU only.

  python tools/make_built_scene.py ggs120 refl_transp variants/c4_built.npz      # config 4 on this repo's tree
  python tools/make_built_scene.py ssc120 refl_transp variants/c4b_built.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rebuilt(sc, cull=True, **tree_kw):
    from ntracer_b200 import bulk
    D, rec = int(sc['dim']), sc['simplex']
    n = rec.shape[0]
    p1 = rec[:, D + 1:2 * D + 1].astype(np.float64)
    E = rec[:, 2 * D + 1:2 * D + 1 + (D - 1) * D].reshape(n, D - 1, D).astype(np.float64)
    pts = np.zeros((n, D, D))
    pts[:, 0] = p1
    for k in range(n):
        pts[k, 1:] = p1[k] - np.linalg.pinv(E[k]).T
    b = bulk.simplex_scene(pts.astype(np.float32), material_ids=sc['simplex_mat'], materials=sc['materials'], batch=4, cull=cull, **tree_kw)
    for k in ('params', 'ambient', 'bg1', 'bg2', 'bg3', 'point_lights', 'global_lights', 'cam_origin', 'cam_axes'):
        b[k] = sc[k]
    return b


def main():
    from tests import fixtures as fx
    name, variant, out = sys.argv[1:4]
    sc, g = fx.load(name)
    if variant and variant != '-':
        sc = fx.variant(sc, g, variant)
    b = rebuilt(sc)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez(out, **b)
    leaves = (b['nodes'][:, 0] & 0x80000000) != 0
    print('%s: %d nodes, %d leaf items, largest leaf %d items (reference tree: %d nodes, %d leaf items)' %
          (out, len(b['nodes']), len(b['leaf_refs']), int(b['nodes'][leaves, 2].max()), len(sc['nodes']), len(sc['leaf_refs'])))


if __name__ == '__main__':
    main()
