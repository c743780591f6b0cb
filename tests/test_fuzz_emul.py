"""Differential fuzzing of the product's per-ray device code (compiled for the host, tests/emul_lib.py) against the
oracle: random scenes mixing 4-lane batches, single simplexes, hypercubes, hyperspheres, opaque / transparent /
reflective materials, shadows and bounce depths, in 3 to 7 dimensions, through the fixed-dimension and the
run-time-dimension instantiations."""
import numpy as np

from tests import emul_lib as el
from tests import fixtures as fx
from tests import oracle_lib as ol


def test_random_mixed_scenes_match_the_oracle():
    w, h = 48, 27
    defined_scenes = solids = transparent_hits = 0
    for seed in range(120):
        dim = 3 + seed % 5
        sc = fx.fuzz_scene(dim, seed)
        a, mask, cnt_o = ol.render_float(sc, w, h, with_mask=True, with_counters=True)
        b, cnt_e = el.render(sc, w, h)
        g, _ = el.render(sc, w, h, generic=True)
        ok = mask == 0                       # pixels where the reference's own lists stay inside their preallocation
        d = np.abs(a - b).max(axis=2)
        assert (d[ok].max() if ok.any() else 0) <= 2e-5, (seed, dim)
        assert np.mean(d > 1e-3) <= 0.01, (seed, dim)          # and even outside that domain nothing drifts far
        assert np.abs(b - g).max() <= 2e-5, (seed, dim)        # fixed- and run-time-dimension code agree
        if ok.all():
            for k in ('primary_rays', 'reflection_rays', 'shadow_rays', 'shaded_hits'):
                assert cnt_o[k] == cnt_e[k], (seed, dim, k)
            defined_scenes += 1
        solids += cnt_o['solid_tests'] > 0
        transparent_hits += cnt_o['reflection_rays'] > 0
    assert defined_scenes >= 40 and solids >= 40 and transparent_hits >= 60      # the corpus exercises what it claims


def test_random_rays_with_skip_primitives_match_the_oracle():
    """KDNode.intersects / occludes hooks on the same corpus: random rays, random `source` primitives and batch lanes
    to skip (tracer.hpp:998-1001), random light distances."""
    rays = 0
    for seed in range(40):
        dim = 3 + seed % 5
        sc = fx.fuzz_scene(dim, seed)
        rng = np.random.RandomState(seed + 999)
        n = 200
        o = np.zeros((n, dim), np.float32)
        o[:, :3] = rng.uniform(-3, 3, (n, 3))
        o[:, 3:] = rng.uniform(-0.05, 0.05, (n, dim - 3))
        target = np.zeros((n, dim), np.float32)
        target[:, :3] = rng.uniform(-1, 1, (n, 3))
        d = (target - o).astype(np.float32)
        refs = np.unique(sc['leaf_refs'])
        skip_ref = rng.choice(refs, size=n).astype(np.uint32)
        skip_lane = np.where((skip_ref >> 30) == 1, rng.randint(-1, 4, size=n), -1).astype(np.int32)
        light = rng.uniform(0.5, 6, n).astype(np.float32)
        oi, od, ont = ol.trace_rays(sc, o, d, skip_ref=skip_ref, skip_lane=skip_lane)
        oocc, _ = ol.occludes_rays(sc, o, d, light, skip_ref, skip_lane)
        for generic in (False, True):
            ids, dist, nt = el.trace_rays(sc, o, d, skip_ref=skip_ref, skip_lane=skip_lane, generic=generic)
            assert np.array_equal(ids, oi) and np.array_equal(nt, ont), (seed, generic)
            assert np.allclose(dist, od, rtol=1e-6, atol=1e-6)
        occ, _ = el.occludes_rays(sc, o, d, light, skip_ref, skip_lane)
        assert np.array_equal(occ, oocc), seed
        rays += n
        assert (oi >= 0).any()
    assert rays == 8000
