"""Pins the oracle (oracle/ntr_oracle.c, the C restatement of the reference's per-pixel path) against golden
vectors produced by the REAL reference (tests/golden/make_fixtures.py), including the reference's own
known-answer test test_kdtree (reference lib/ntracer/tests/test.py:302-363).  CPU only."""
import numpy as np
import pytest

from ntracer_b200 import _capi
from tests import fixtures as fx
from tests import oracle_lib as ol


def test_reference_kdtree_known_answer():
    sc, g = fx.load('kdtree_kat')
    ids, dist, nt = ol.trace_rays(sc, g['origin'][None], g['direction'][None])
    assert ids[0] == int(g['expected_id'])          # primitive index 4 of the reference test
    assert nt[0] == 0                               # exactly one hit
    assert dist[0] == pytest.approx(float(g['expected_dist']), rel=1e-5)
    ids, dist, nt = ol.trace_rays(sc, g['fan_origins'], g['fan_dirs'])
    assert np.array_equal(ids, g['fan_ids'])
    assert np.allclose(dist, g['fan_dists'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('dim', [3, 4, 6, 9])
def test_box_scene_bytes(dim):
    sc, g = fx.load('box%d' % dim)
    w, h = [int(v) for v in g['size']]
    fmt = _capi.make_image_format(w, h, _capi.RGB8)
    img = ol.render_packed(sc, fmt)
    d = np.abs(img.astype(np.int32) - g['packed'].astype(np.int32))
    # the reference is compiled with -ffast-math (rsqrt-based unit()): bytes at exact .5 boundaries may flip
    assert d.max() <= 1
    assert np.count_nonzero(d) <= 1e-5 * d.size + 2
    for (x, y), c in zip(g['points'], g['colors']):
        assert np.allclose(ol.calculate_color(sc, int(x), int(y), w, h), c, atol=2e-7)


def test_survey_golden_values():
    # values recorded in SURVEY.md section 8(c) from the compiled reference
    sc, g = fx.load('box4')
    assert np.allclose(ol.calculate_color(sc, 100, 200, 640, 480), (0, 0.27875945, 0.27875945), atol=1e-7)
    assert np.allclose(ol.screen_coord_to_ray(4, sc['cam_axes'], 100, 200, 640, 480, 0.8),
                       (-0.27875945, 0.05068354, 0.95902264, 0), atol=1e-7)


def test_pack_formats():
    sc, g = fx.load('pack')
    w, h = [int(v) for v in g['size']]
    for n in g['names']:
        n = str(n)
        ch = [(int(c[0]), c[1], c[2], c[3], c[4], bool(c[5])) for c in g['fmt_' + n]]
        pitch, rev = [int(v) for v in g['opt_' + n]]
        fmt = _capi.make_image_format(w, h, ch, pitch, bool(rev))
        out = ol.pack(fmt, g['float']).reshape(h, fmt.pitch)[:, :w * fmt.bytes_per_pixel]
        ref = g['out_' + n].reshape(h, fmt.pitch)[:, :w * fmt.bytes_per_pixel]
        if n == 'wide':
            # its float channel mixes r,g,b with non-trivial weights: the reference's -ffast-math summation
            # order may differ in the last ulp -> only the two low bytes of that channel may differ
            bad = (out != ref).reshape(h, w, fmt.bytes_per_pixel)
            assert not bad[:, :, :14].any()
        else:
            assert np.array_equal(out, ref), n


@pytest.mark.parametrize('name', ['cell120', 'ggs120', 'ssc120'])
def test_polytope_variants(name):
    sc, g = fx.load(name)
    w, h = [int(v) for v in g['size']]
    col = fx.center_column_mask(w, h)
    for v in g['variants']:
        v = str(v)
        img, mask = ol.render_float(fx.variant(sc, g, v), w, h, with_mask=True)
        gold = g['v_%s_float' % v]
        # Where the reference is defined, the restatement must agree (tolerance of BASELINE.json).  Pixels where a
        # reference quick_list outgrew its preallocation are undefined behaviour in the reference itself
        # (tracer.hpp:670-680: uninitialised mailbox slots -> primitives skipped at random, seen as holes in its
        # {5/2,3,3} images); they are only held to a loose bound.
        undefined = (mask & 3) != 0
        bad_defined, _ = fx.lsb_stats(img, gold, exclude=col | undefined)
        bad_all, _ = fx.lsb_stats(img, gold, exclude=col)
        assert bad_defined <= 0.001, (name, v, bad_defined)
        assert bad_all <= 0.03, (name, v, bad_all)
    ids, dist = ol.primary_hit_ids(sc, w, h)
    agree, ties = fx.id_agreement(ids, g['ids'], dist, g['dist'])
    assert agree >= 0.9999
    raw, _ = fx.id_agreement(ids, g['ids'])
    assert raw >= 0.999


def test_cell120_occludes():
    sc, g = fx.load('cell120')
    occ, nt = ol.occludes_rays(sc, g['occ_origins'], g['occ_dirs'], g['occ_dist'], g['occ_skip_ref'], g['occ_skip_lane'])
    assert np.mean(occ == g['occ_result']) >= 0.995


@pytest.mark.parametrize('name', ['solids6', 'soup9'])
def test_solids_and_generic_dimension(name):
    sc, g = fx.load(name)
    w, h = [int(v) for v in g['size']]
    img = ol.render_float(sc, w, h)
    assert fx.lsb_stats(img, g['float'])[0] <= 0.001
    ids, dist = ol.primary_hit_ids(sc, w, h)
    assert fx.id_agreement(ids, g['ids'], dist, g['dist'])[0] >= 0.9999


def test_mixed_transparent_scene():
    sc, g = fx.load('mixed3')
    w, h = [int(v) for v in g['size']]
    ids, dist, nt = ol.trace_rays(sc, g['ray_origins'], g['ray_dirs'])
    assert np.array_equal(ids, g['ray_ids'])
    assert np.array_equal(nt, g['ray_ntrans'])
    assert np.allclose(dist, g['ray_dists'], rtol=1e-5, atol=1e-5)
    img, mask = ol.render_float(sc, w, h, with_mask=True)
    # pixels whose opaque hit had its normal overwritten by a later transparent hit (reference quirk, DESIGN.md
    # Q12; oracle mask bit 2) shoot secondary rays from ON a transparent surface: their t ~ 0 self-hits are rounding
    # noise.  Every other pixel agrees with the reference's frame.
    assert (mask & 3).max() == 0 and 0.02 <= np.mean(mask == 4) <= 0.1
    assert fx.lsb_stats(img, g['float'], exclude=mask != 0)[0] == 0
    assert fx.lsb_stats(img, g['float'])[0] <= 0.015
