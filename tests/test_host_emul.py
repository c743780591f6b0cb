"""Runs the product's per-ray device code (ntracer_b200/csrc/trace_core.cuh: the stack-machine traversal,
primitive tests, shading and pixel packing that the CUDA kernels inline) compiled for the HOST, and checks it
against the oracle and the golden vectors.  This is how the kernel logic is exercised in the CPU-only tier;
the `-m gpu` tests run the real kernels through the C ABI."""
import numpy as np
import pytest

from ntracer_b200 import _capi
from tests import emul_lib as el
from tests import fixtures as fx
from tests import oracle_lib as ol


@pytest.mark.parametrize('name,variants', [
    ('cell120', ['camlight', 'shadows', 'refl', 'refl_transp', 'transp', 'depth1_spec']),
    ('ggs120', ['refl', 'refl_transp']),
    ('ssc120', ['camlight_lights', 'refl', 'refl_transp']),
])
def test_polytope_matches_oracle(name, variants):
    sc, g = fx.load(name)
    w, h = 96, 54
    for v in variants:
        s2 = fx.variant(sc, g, v)
        a, mask, cnt_o = ol.render_float(s2, w, h, with_mask=True, with_counters=True)
        b, cnt_e = el.render(s2, w, h)
        # (giant leaves included: scenes whose leaves overrun the bounded mailbox table get the exact per-thread bitset,
        # trace_core.cuh: MailboxStore, which is the oracle's unbounded list -- same tests, same counters, same picture,
        # also on the half of a {5/2,3,3} frame where the reference itself is undefined)
        assert np.abs(a - b).max() <= 2e-6, (name, v)
        for k in ('primary_rays', 'reflection_rays', 'shadow_rays', 'node_steps', 'shaded_hits'):
            assert cnt_o[k] == cnt_e[k], (name, v, k)
        if name != 'cell120' and 'transp' in v:
            assert cnt_o['simplex_tests'] == cnt_e['simplex_tests'], (name, v)     # the exact mailbox tests what the oracle tests


@pytest.mark.parametrize('name', ['solids6', 'mixed3', 'soup9', 'box4', 'box9'])
def test_scene_matches_oracle_and_golden(name):
    sc, g = fx.load(name)
    w, h = [int(v) for v in g['size']]
    a = ol.render_float(sc, w, h)
    b, ids, dist, cnt = el.render(sc, w, h, want_ids=True)
    assert np.abs(a - b).max() <= 2e-6
    if 'ids' in g:
        assert fx.id_agreement(ids, g['ids'], dist, g['dist'])[0] >= 0.9999


def test_runtime_dimension_variant_equals_fixed():
    sc, g = fx.load('cell120')
    a, _ = el.render(sc, 64, 36)
    b, _ = el.render(sc, 64, 36, generic=True)
    assert np.abs(a - b).max() <= 2e-6
    sc, g = fx.load('solids6')
    a, _ = el.render(sc, 64, 36)
    b, _ = el.render(sc, 64, 36, generic=True)
    assert np.abs(a - b).max() <= 2e-6


def test_kdtree_known_answer_and_rays():
    sc, g = fx.load('kdtree_kat')
    ids, dist, nt = el.trace_rays(sc, g['origin'][None], g['direction'][None])
    assert ids[0] == int(g['expected_id']) and nt[0] == 0
    ids, dist, nt = el.trace_rays(sc, g['fan_origins'], g['fan_dirs'])
    assert np.array_equal(ids, g['fan_ids'])
    sc, g = fx.load('mixed3')
    for generic in (False, True):
        ids, dist, nt = el.trace_rays(sc, g['ray_origins'], g['ray_dirs'], generic=generic)
        assert np.array_equal(ids, g['ray_ids'])
        assert np.array_equal(nt, g['ray_ntrans'])
        assert np.allclose(dist, g['ray_dists'], rtol=1e-5, atol=1e-5)
    sc, g = fx.load('cell120')
    occ, nt = el.occludes_rays(sc, g['occ_origins'], g['occ_dirs'], g['occ_dist'], g['occ_skip_ref'], g['occ_skip_lane'])
    occ_o, _ = ol.occludes_rays(sc, g['occ_origins'], g['occ_dirs'], g['occ_dist'], g['occ_skip_ref'], g['occ_skip_lane'])
    assert np.array_equal(occ, occ_o)
    assert np.mean(occ == g['occ_result']) >= 0.995


def test_pack_pixel_formats():
    sc, g = fx.load('pack')
    w, h = [int(v) for v in g['size']]
    for n in g['names']:
        n = str(n)
        ch = [(int(c[0]), c[1], c[2], c[3], c[4], bool(c[5])) for c in g['fmt_' + n]]
        pitch, rev = [int(v) for v in g['opt_' + n]]
        fmt = _capi.make_image_format(w, h, ch, pitch, bool(rev))
        assert np.array_equal(el.pack(fmt, g['float']), ol.pack(fmt, g['float'])), n


@pytest.mark.parametrize('dim', [3, 5, 7, 8, 9, 10, 12])
def test_batched_soup_every_fixed_dimension_and_generic(dim):
    """batch_test / simplex_single / shading / shadows / one reflection pass of every fixed-dimension instantiation
    (3..10) and of the run-time-dimension one, on a synthetic scene with 4-lane batches (fixtures.batched_soup)."""
    sc = fx.batched_soup(dim, 40)
    w, h = 64, 36
    a, cnt_o = ol.render_float(sc, w, h, with_counters=True)
    assert cnt_o['shadow_rays'] > 500 and cnt_o['reflection_rays'] > 200          # the scene exercises both
    b, ids, dist, cnt = el.render(sc, w, h, want_ids=True)
    assert np.abs(a - b).max() <= 2e-6
    oids, odist = ol.primary_hit_ids(sc, w, h)
    assert fx.id_agreement(ids, oids, dist, odist)[0] >= 0.9999
    for k in ('primary_rays', 'reflection_rays', 'shadow_rays', 'shaded_hits'):
        assert cnt_o[k] == cnt[k], (dim, k)
    g, _ = el.render(sc, w, h, generic=True)
    assert np.abs(a - g).max() <= 2e-6


def test_truncated_hit_lists_are_reported_and_too_deep_reflection_is_rejected():
    """Limits that would change the picture are not silent: rays with more transparent layers than the kernels' lists hold
    (16; the reference itself is undefined beyond 10) are counted in ntr_counters.truncated_hit_lists."""
    w, h = 24, 18
    sc = fx.stacked_layers(12)
    a, mask, cnt_o = ol.render_float(sc, w, h, with_mask=True, with_counters=True)
    b, cnt_e = el.render(sc, w, h)
    assert cnt_e['truncated_hit_lists'] == 0 and np.abs(a - b).max() <= 2e-6      # 12 layers: beyond the reference's 10, fine here
    assert (mask & 1).any()                                                        # ... and the oracle says so
    sc = fx.stacked_layers(20)
    b, cnt_e = el.render(sc, w, h)
    assert cnt_e['truncated_hit_lists'] > 0


def test_tiny_and_ragged_frames_and_the_empty_scene():
    """CPU twin of the GPU edge-case test: frames of one pixel / smaller than a block / ending inside a block, and a
    CompositeScene without primitives (no tree: background gradient only), per-ray device code against the oracle."""
    sc, g = fx.load('cell120')
    sc = fx.variant(sc, g, 'shadows')
    for w, h in ((1, 1), (3, 2), (9, 5), (33, 31), (257, 3)):
        a = ol.render_float(sc, w, h)
        b, _ = el.render(sc, w, h)
        assert a.shape == b.shape == (h, w, 3) and np.abs(a - b).max() <= 2e-6
    e = dict(sc)
    e['nodes'] = np.zeros((0, 4), np.uint32)
    e['leaf_refs'] = np.zeros(0, np.uint32)
    e['root'] = np.int64(0xFFFFFFFF)
    e['simplex'] = np.zeros((0, sc['simplex'].shape[1]), np.float32)
    e['simplex_mat'] = np.zeros(0, np.int32)
    a = ol.render_float(e, 32, 18)
    b, ids, dist, cnt = el.render(e, 32, 18, want_ids=True)
    assert np.abs(a - b).max() <= 2e-6 and (ids == -1).all()
    assert cnt['simplex_tests'] == 0 and cnt['shaded_hits'] == 0
    assert np.ptp(a.reshape(-1, 3), axis=0).max() > 0.05             # the gradient, not a constant
