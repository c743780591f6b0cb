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
