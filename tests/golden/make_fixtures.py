#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref, built from /root/reference by
oracle/build_ref.sh).  Runs only in the build container; the fixtures it writes are committed so that
the oracle (oracle/ntr_oracle.c) and the CUDA path can be pinned on machines without the reference.

  python tests/golden/make_fixtures.py            # everything
  python tests/golden/make_fixtures.py box cell120   # a subset

Every fixture holds a flat scene (format: oracle/ref_bridge.py) plus `g_*` arrays = outputs of the
reference itself: Scene.calculate_color (float RGB), KDNode.intersects / occludes (hit ids),
BlockingRenderer.render (packed frames).
"""
import hashlib
import os
import random
import sys

import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)
import ref_bridge as rb  # noqa: E402

ntracer = rb.load_reference()
RGB8 = [(8, 1, 0, 0), (8, 0, 1, 0), (8, 0, 0, 1)]


def ref_format(w, h, channels, pitch=0, reversed=False):
    return ntracer.ImageFormat(w, h, [ntracer.Channel(*c) for c in channels], pitch, reversed)


def ref_float(scene, w, h):
    out = np.zeros((h, w, 3), np.float32)
    for y in range(h):
        for x in range(w):
            c = scene.calculate_color(x, y, w, h)
            out[y, x] = (c.r, c.g, c.b)
    return out


def ref_packed(scene, w, h, channels=RGB8, pitch=0, reversed=False, fill=0xAB):
    fmt = ref_format(w, h, channels, pitch, reversed)
    buf = bytearray([fill]) * (fmt.pitch * h)
    renderer = ntracer.BlockingRenderer()
    time.sleep(0.5)      # reference start-up race: workers must be waiting before the first job (see bench.reference_arm)
    assert renderer.render(buf, fmt, scene)
    return np.frombuffer(bytes(buf), dtype=np.uint8).copy()


def ref_primary_ids(nt, scene, sc, w, h):
    """scene.root.intersects for every primary ray -> flat primitive ids (-1 = miss) and distances."""
    cam = scene.get_camera()
    ids = np.full((h, w), -1, np.int32)
    dist = np.zeros((h, w), np.float32)
    item_ref = sc['_item_ref']
    for y in range(h):
        for x in range(w):
            d = nt.screen_coord_to_ray(cam, x, y, w, h, scene.fov)
            hits = scene.root.intersects(cam.origin, d)
            if hits and (hits[-1].primitive.material.opacity >= 1 if hits[-1].batch_index < 0
                         else hits[-1].primitive[hits[-1].batch_index].material.opacity >= 1):
                hit = hits[-1]
                ids[y, x] = rb.flat_prim_id(sc, item_ref[id(hit.primitive)], hit.batch_index)
                dist[y, x] = hit.dist
    return ids, dist


def save(name, sc, **golden):
    path = os.path.join(HERE, name + '.npz')
    data = rb.strip_private(sc)
    data.update({'g_' + k: v for k, v in golden.items()})
    np.savez_compressed(path, **data)
    print('wrote %s (%.1f KiB)' % (path, os.path.getsize(path) / 1024))


def variant_arrays(sc):
    """The arrays that change between variants of one scene (no geometry)."""
    return {k: sc[k] for k in ('params', 'materials', 'point_lights', 'global_lights', 'ambient', 'bg1', 'bg2', 'bg3')}


# ---------------------------------------------------------------------------------------------------
def make_box():
    # config 1: 4-D tesseract BoxScene, scripts/hypercube.py:309,356-361, 640x480
    for dim, (w, h) in ((4, (640, 480)), (6, (480, 270)), (3, (160, 120)), (9, (96, 64))):
        nt = ntracer.NTracer(dim)
        scene = nt.BoxScene()
        cam = nt.Camera()
        cam.translate(nt.Vector.axis(2, -5))
        if dim != 4:   # an off-axis view as well
            cam.transform(nt.Matrix.rotation(nt.Vector.axis(0), nt.Vector.axis(2), 0.3))
            cam.transform(nt.Matrix.rotation(nt.Vector.axis(1), nt.Vector.axis(dim - 1), 0.2))
            cam.normalize()
            cam.origin = cam.axes[2] * -5
        scene.set_camera(cam)
        sc = rb.export_scene(nt, scene)
        pts = [(w // 2, h // 2), (0, 0), (100 % w, 200 % h), (w - 140, 50), (w // 2 - 60, h // 2 + 45)]
        cols = np.array([list(scene.calculate_color(x, y, w, h)) for x, y in pts], np.float32)
        packed = ref_packed(scene, w, h)
        save('box%d' % dim, sc, size=np.array([w, h]), points=np.array(pts), colors=cols, packed=packed,
             md5=np.frombuffer(hashlib.md5(packed.tobytes()).digest(), dtype=np.uint8))


def make_pack():
    # image-format coverage (process_pixel, render.cpp:396-466): odd bit sizes, >64-bit pixels, float channels,
    # reversed byte order, pitch padding (padding bytes must stay untouched)
    nt = ntracer.NTracer(4)
    scene = nt.BoxScene()
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -5))
    scene.set_camera(cam)
    sc = rb.export_scene(nt, scene)
    w, h = 70, 37
    formats = {
        'rgb565': dict(channels=[(5, 1, 0, 0), (6, 0, 1, 0), (5, 0, 0, 1)]),
        'rgb16': dict(channels=[(16, 1, 0, 0), (16, 0, 1, 0), (16, 0, 0, 1)]),
        'bgrx_rev': dict(channels=[(8, 0, 0, 0, 1.0), (8, 1, 0, 0), (8, 0, 1, 0), (8, 0, 0, 1)], reversed=True),
        'gray1': dict(channels=[(1, 0.3, 0.6, 0.1)]),
        'pitch': dict(channels=RGB8, pitch=70 * 3 + 7),
        'float3': dict(channels=[(32, 1, 0, 0, 0, True), (32, 0, 1, 0, 0, True), (32, 0, 0, 1, 0, True)]),
        'wide': dict(channels=[(31, 1, 0, 0), (31, 0, 1, 0), (31, 0, 0, 1), (32, 0.2, 0.7, 0.1, 0, True), (3, 0, 0, 0, 0.5)]),
        'odd': dict(channels=[(3, 1, 0, 0), (11, 0, 1, 0), (7, 0, 0, 1), (2, 0.5, 0.5, 0)], reversed=True, pitch=70 * 3 + 1),
        'straddle': dict(channels=[(30, 1, 0, 0), (30, 0, 1, 0), (30, 0, 0, 1), (30, 0.5, 0.5, 0)]),
    }
    golden = {'size': np.array([w, h]), 'float': ref_float(scene, w, h)}
    names = []
    for name, f in formats.items():
        ch = f['channels']
        golden['fmt_' + name] = np.array([list(c) + [0] * (6 - len(c)) for c in ch], np.float64)
        golden['opt_' + name] = np.array([f.get('pitch', 0), int(f.get('reversed', False))])
        golden['out_' + name] = ref_packed(scene, w, h, ch, f.get('pitch', 0), f.get('reversed', False))
        names.append(name)
    golden['names'] = np.array(names)
    save('pack', sc, **golden)


def make_kdtree_kat():
    # the reference's own known-answer test: lib/ntracer/tests/test.py:302-363 (test_kdtree)
    nt = ntracer.NTracer(3)
    mat = ntracer.Material((1, 1, 1))
    T = nt.Triangle
    prims = [
        T((-1.1755770444869995, 0.3819499611854553, -1.6180520057678223), (1.7082732915878296, -2.3512351512908936, 1.4531432390213013),
          [(-0.615524172782898, -0.3236003816127777, 0.19999605417251587), (0.49796950817108154, 0.0381958931684494, -0.5235964059829712)], mat),
        T((-1.1755770444869995, 0.3819499611854553, -1.6180520057678223), (1.0557708740234375, -1.4531433582305908, 0.8980922102928162),
          [(-0.8057316541671753, -0.06180214881896973, 0.8471965789794922), (0.19020742177963257, -0.2617982029914856, -0.6472004652023315)], mat),
        T((0.7265498042106628, 0.9999955296516418, 1.6180428266525269), (0, 1.7961481809616089, 0.8980742692947388),
          [(-1.1135050058364868, -0.1618017703294754, 0.32360348105430603), (0.6881839036941528, -0.09999901801347733, 0.19999800622463226)], mat),
        T((0.7265498042106628, 0.9999955296516418, 1.6180428266525269), (0, 2.90622878074646, 1.4531147480010986),
          [(-0.4253210127353668, -0.26180076599121094, 0.5236014127731323), (0.6881839036941528, 0.09999898821115494, -0.1999979317188263)], mat),
        T((1.9021340608596802, 0.618022620677948, -0.3819592595100403), (-1.055770754814148, -1.4531432390213013, 0.8980920910835266),
          [(-0.30776214599609375, -0.42359834909439087, -1.0471925735473633), (0.4979696571826935, -0.038195837289094925, 0.5235962867736816)], mat),
        T((1.9021340608596802, 0.618022620677948, -0.3819592595100403), (-1.7082730531692505, -2.3512353897094727, 1.4531434774398804),
          [(0.19020749628543854, -0.4617941677570343, -0.5235962271690369), (0.19020745158195496, 0.2617981433868408, 0.6472005844116211)], mat)]
    scene = nt.CompositeScene(
        nt.AABB((-1.710653305053711e-05, 0.618022620677948, -0.3819774389266968), (0.7265291213989258, 2.000016689300537, 0.3819882869720459)),
        nt.KDBranch(1, 2.0000057220458984,
                    nt.KDBranch(1, 0.9999955296516418, None,
                                nt.KDLeaf([prims[4], prims[5], prims[2], prims[3], prims[1], prims[0]])),
                    nt.KDLeaf([prims[4], prims[5], prims[1], prims[0]])))
    scene.set_fov(0.8)
    sc = rb.export_scene(nt, scene)
    origin = (4.917067527770996, 2.508934497833252, -4.304379940032959)
    direction = (-0.7135500907897949, -0.1356230527162552, 0.6873518228530884)
    hits = scene.root.intersects(origin, direction)
    assert len(hits) == 1 and prims.index(hits[0].primitive) == 4 and hits[0].batch_index == -1
    item_ref = sc['_item_ref']
    # flat ids follow export order, not `prims` order: record the mapping of the expected answer
    expected = rb.flat_prim_id(sc, item_ref[id(prims[4])], -1)
    # plus a fan of rays for good measure
    rng = random.Random(7)
    origins, dirs, ids, dists = [], [], [], []
    for i in range(400):
        o = [rng.uniform(-6, 6) for _ in range(3)]
        tgt = [rng.uniform(-1, 1.5), rng.uniform(0, 2.5), rng.uniform(-1, 1.5)]
        d = nt.Vector(*[t - a for t, a in zip(tgt, o)]).unit()
        hs = scene.root.intersects(tuple(o), d)
        origins.append(o)
        dirs.append(list(d))
        ids.append(rb.flat_prim_id(sc, item_ref[id(hs[-1].primitive)], hs[-1].batch_index) if hs else -1)
        dists.append(hs[-1].dist if hs else 0)
    save('kdtree_kat', sc, origin=np.array(origin, np.float32), direction=np.array(direction, np.float32),
         expected_id=np.int64(expected), expected_dist=np.float32(hits[0].dist),
         fan_origins=np.array(origins, np.float32), fan_dirs=np.array(dirs, np.float32), fan_ids=np.array(ids, np.int32),
         fan_dists=np.array(dists, np.float32))


def first_material(nt, scene):
    n = scene.root
    while not isinstance(n, nt.KDLeaf):
        n = n.left if n.left is not None else n.right
    it = n[0]
    return it[0].material if isinstance(it, nt.TriangleBatch) else it.material


def add_c2_lights(nt, scene):
    # the lights polytope.py never adds, fixed here for both sides
    scene.add_light(nt.PointLight(nt.Vector.axis(1, 8) + nt.Vector.axis(2, -8), (60, 60, 60)))
    scene.add_light(nt.GlobalLight(nt.Vector.axis(1, -1), (0.4, 0.4, 0.4)))


def make_cell120():
    # config 2: {5,3,3} 120-cell via the reference's scripts/polytope.py + build_composite_scene
    nt, scene, cam = rb.polytope_scene('5 3 3')
    w, h = 192, 108
    golden = {'size': np.array([w, h])}
    sc = rb.export_scene(nt, scene)
    base = sc
    golden['v_camlight_float'] = ref_float(scene, w, h)
    ids, dist = ref_primary_ids(nt, scene, sc, w, h)
    golden['ids'], golden['dist'] = ids, dist
    for k, v in variant_arrays(sc).items():
        golden['v_camlight_' + k] = v
    add_c2_lights(nt, scene)
    mat = first_material(nt, scene)
    variants = [
        ('lights', dict(shadows=False)),
        ('shadows', dict(shadows=True)),                                   # = config 2
        ('refl', dict(shadows=True, reflectivity=0.3)),
        ('refl_transp', dict(shadows=True, reflectivity=0.3, opacity=0.5)),
        ('transp', dict(shadows=True, opacity=0.5)),
        ('transp_noshadow', dict(shadows=False, opacity=0.5)),
        ('depth1_spec', dict(shadows=True, reflectivity=0.5, depth=1, spec_exp=20.0, spec_int=0.7, ambient=(0.05, 0.02, 0.1),
                             bg=((0.2, 0.3, 0.9), (0.9, 0.9, 0.9), (0.1, 0.5, 0.1), 0))),
    ]
    for name, v in variants:
        scene.set_shadows(v.get('shadows', False))
        mat.reflectivity = v.get('reflectivity', 0.0)
        mat.opacity = v.get('opacity', 1.0)
        mat.specular_exp = v.get('spec_exp', 8.0)
        mat.specular_intensity = v.get('spec_int', 1.0)
        scene.set_max_reflect_depth(v.get('depth', 4))
        scene.set_ambient_color(v.get('ambient', (0, 0, 0)))
        bg = v.get('bg', ((1, 1, 1), (0, 0, 0), (0, 1, 1), 1))
        scene.set_background(*bg)
        s2 = rb.export_scene(nt, scene)
        assert np.array_equal(s2['nodes'], base['nodes']) and np.array_equal(s2['simplex'], base['simplex'])
        golden['v_%s_float' % name] = ref_float(scene, w, h)
        for k, a in variant_arrays(s2).items():
            golden['v_%s_%s' % (name, k)] = a
        if name == 'shadows':
            golden['v_shadows_packed'] = ref_packed(scene, w, h)
    golden['variants'] = np.array(['camlight'] + [n for n, _ in variants])
    # the committed scene file carries the config-2 state (lights, shadows on, opaque)
    scene.set_shadows(True)
    mat.reflectivity, mat.opacity, mat.specular_exp, mat.specular_intensity = 0.0, 1.0, 8.0, 1.0
    scene.set_max_reflect_depth(4)
    scene.set_ambient_color((0, 0, 0))
    scene.set_background((1, 1, 1), (0, 0, 0), (0, 1, 1), 1)
    sc = rb.export_scene(nt, scene)
    # shadow-ray oracle hook: KDNode.occludes for rays leaving the visible surface points towards the global light
    item_ref = sc['_item_ref']
    cam_o = scene.get_camera().origin
    rng = random.Random(11)
    o_list, d_list, dist_list, sr, sl, occ = [], [], [], [], [], []
    for i in range(600):
        x, y = rng.randrange(w), rng.randrange(h)
        d = nt.screen_coord_to_ray(scene.get_camera(), x, y, w, h, scene.fov)
        hs = scene.root.intersects(cam_o, d)
        if not hs:
            continue
        hit = hs[-1]
        ldir = nt.Vector.axis(1, 1) if i % 2 else (hit.origin - (nt.Vector.axis(1, 8) + nt.Vector.axis(2, -8))).unit()
        ld = 3.4028234663852886e38 if i % 2 else abs(hit.origin - (nt.Vector.axis(1, 8) + nt.Vector.axis(2, -8)))
        r = scene.root.occludes(hit.origin, ldir, ld, source=hit.primitive, batch_index=hit.batch_index)
        o_list.append(list(hit.origin)); d_list.append(list(ldir)); dist_list.append(ld)
        sr.append(item_ref[id(hit.primitive)]); sl.append(hit.batch_index); occ.append(int(r[0]))
    golden.update(occ_origins=np.array(o_list, np.float32), occ_dirs=np.array(d_list, np.float32),
                  occ_dist=np.array(dist_list, np.float32), occ_skip_ref=np.array(sr, np.uint32),
                  occ_skip_lane=np.array(sl, np.int32), occ_result=np.array(occ, np.int32))
    save('cell120', sc, **golden)


def make_star(fixture='ggs120', schlafli='5/2 3 3'):
    # config 4: great grand stellated 120-cell {5/2,3,3}; the same recipe serves
    # {5/2,5,3} (small stellated 120-cell), the symbol BASELINE.json's configs[3] literally names
    nt, scene, cam = rb.polytope_scene(schlafli)
    add_c2_lights(nt, scene)
    mat = first_material(nt, scene)
    w, h = 128, 72
    golden = {'size': np.array([w, h])}
    sc = rb.export_scene(nt, scene)
    ids, dist = ref_primary_ids(nt, scene, sc, w, h)
    golden['ids'], golden['dist'] = ids, dist
    variants = [('camlight_lights', dict(shadows=False)), ('refl', dict(shadows=True, reflectivity=0.3))]
    for name, v in variants:
        scene.set_shadows(v.get('shadows', False))
        mat.reflectivity = v.get('reflectivity', 0.0)
        s2 = rb.export_scene(nt, scene)
        golden['v_%s_float' % name] = ref_float(scene, w, h)
        for k, a in variant_arrays(s2).items():
            golden['v_%s_%s' % (name, k)] = a
    sc = rb.export_scene(nt, scene)     # committed state: lights, shadows, reflectivity 0.3, depth 4
    # config 4 proper: "reflections + transparency".  Every 10th of the 120 cells (cells = clusters of coplanar
    # simplexes) gets opacity 0.5; with 12 transparent cells no ray collects more than 10 transparent hits, i.e. the
    # reference's quick_list stays inside its preallocation.  The materials are assigned in the
    # flat scene and the scene is rebuilt inside the reference through its public constructors.
    S = sc['simplex']
    nrm = np.linalg.norm(S[:, :4], axis=1, keepdims=True)
    key = np.concatenate([S[:, :4] / nrm, S[:, 4:5] / nrm], axis=1)
    key *= np.sign(key[np.arange(len(key)), np.argmax(np.abs(key[:, :4]), axis=1)])[:, None]
    _, cell = np.unique(np.round(key, 3), axis=0, return_inverse=True)
    assert cell.max() == 119
    s2 = rb.strip_private(sc)
    s2['simplex_mat'] = np.where(cell % 10 == 0, 1, 0).astype(np.int32)
    m0 = sc['materials'][0].copy()
    m1 = m0.copy()
    m1[6] = 0.5
    s2['materials'] = np.stack([m0, m1])
    nt2, scene2, prims2 = rb.import_scene(s2)
    golden['v_refl_transp_float'] = ref_float(scene2, w, h)
    for k, a in variant_arrays(s2).items():
        golden['v_refl_transp_' + k] = a
    golden['v_refl_transp_simplex_mat'] = s2['simplex_mat']
    golden['variants'] = np.array([n for n, _ in variants] + ['refl_transp'])
    save(fixture, sc, **golden)


def make_ggs120():
    make_star('ggs120', '5/2 3 3')


def make_ssc120():
    make_star('ssc120', '5/2 5 3')


def rot(nt, i, j, theta):
    return nt.Matrix.rotation(nt.Vector.axis(i), nt.Vector.axis(j), theta)


def make_solids6():
    # config 3 stand-in: 6-D solids, reflections depth 4, one point light
    nt = ntracer.NTracer(6)
    M = ntracer.Material
    cube_o = rot(nt, 0, 2, 0.5) * rot(nt, 1, 3, 0.4) * rot(nt, 0, 4, 0.3) * rot(nt, 2, 5, 0.2)
    floor_o = nt.Matrix.scale(nt.Vector(10, 1, 10, 10, 10, 10))
    protos = [
        nt.SolidPrototype(ntracer.CUBE, nt.Vector(0, 0, 0, 0, 0, 0), cube_o, M((1, 0.5, 0.5), 1, 0.5)),
        nt.SolidPrototype(ntracer.CUBE, nt.Vector(0, -3, 0, 0, 0, 0), floor_o, M((0.5, 0.5, 1), 1, 0.5)),
        nt.SolidPrototype(ntracer.SPHERE, nt.Vector(0, 0, 3, 0, 0, 0), nt.Matrix.identity(), M((0.5, 1, 0.5), 1, 0.5)),
        nt.SolidPrototype(ntracer.SPHERE, nt.Vector(2.5, 0, 0, 0, 0, 0), nt.Matrix.scale(0.8), M((1, 1, 0.3), 0.6, 0.2)),
    ]
    scene = nt.build_composite_scene(protos)
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -7) + nt.Vector.axis(1, 1.0))
    cam.transform(rot(nt, 2, 3, 0.15))
    cam.transform(rot(nt, 0, 4, 0.1))
    cam.transform(rot(nt, 1, 5, 0.05))
    cam.normalize()
    scene.set_camera(cam)
    scene.add_light(nt.PointLight(nt.Vector(3, 6, -6, 1, 0, 0), (400, 400, 400)))
    scene.set_shadows(True)
    w, h = 160, 90
    sc = rb.export_scene(nt, scene)
    ids, dist = ref_primary_ids(nt, scene, sc, w, h)
    save('solids6', sc, size=np.array([w, h]), float=ref_float(scene, w, h), ids=ids, dist=dist,
         packed=ref_packed(scene, w, h))


def make_mixed3():
    # 3-D hand-built tree mixing opaque + transparent triangles, a batch, a cube and a sphere: exercises the
    # transparent-hit list, the mailbox, trim_intersections and the `source` skip rules in their defined regime
    nt = ntracer.NTracer(3)
    M = ntracer.Material
    opaque = M((0.9, 0.9, 0.2), 1, 0.4)
    glass = M((0.2, 0.6, 1.0), 0.4, 0.0)
    tint = M((1.0, 0.3, 0.3), 0.7, 0.3, 0.5, 12.0, (0.9, 1, 0.8))
    rng = random.Random(5)

    def tri(cx, cy, cz, s, m):
        pts = [nt.Vector(cx + rng.uniform(-s, s), cy + rng.uniform(-s, s), cz + rng.uniform(-s, s)) for _ in range(3)]
        return nt.TrianglePrototype(pts, m)

    protos = []
    for i in range(60):
        m = (opaque, glass, tint)[i % 3]
        protos.append(tri(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3), 1.2, m))
    protos.append(nt.SolidPrototype(ntracer.CUBE, nt.Vector(0, 0, 0), rot(nt, 0, 1, 0.3) * rot(nt, 1, 2, 0.2), glass))
    protos.append(nt.SolidPrototype(ntracer.SPHERE, nt.Vector(1.5, 1, -1), nt.Matrix.scale(0.7), opaque))
    protos.append(nt.SolidPrototype(ntracer.CUBE, nt.Vector(0, -4.5, 0), nt.Matrix.scale(nt.Vector(6, 1, 6)), M((0.6, 0.6, 0.6), 1, 0.3)))
    scene = nt.build_composite_scene(protos)
    cam = nt.Camera()
    cam.translate(nt.Vector(0.5, 1.0, -10))
    scene.set_camera(cam)
    scene.add_light(nt.PointLight(nt.Vector(4, 8, -6), (150, 150, 150)))
    scene.add_light(nt.GlobalLight(nt.Vector(0.3, -1, 0.2), (0.5, 0.5, 0.4)))
    scene.set_shadows(True)
    scene.set_max_reflect_depth(3)
    w, h = 128, 96
    sc = rb.export_scene(nt, scene)
    item_ref = sc['_item_ref']
    # KDNode.intersects on random rays: opaque id + surviving transparent count
    origins, dirs, ids, dists, ntr = [], [], [], [], []
    for i in range(1500):
        o = [rng.uniform(-8, 8) for _ in range(3)]
        tgt = [rng.uniform(-3, 3) for _ in range(3)]
        d = nt.Vector(*[t - a for t, a in zip(tgt, o)]).unit()
        hs = scene.root.intersects(tuple(o), d)
        opq = None
        if hs:
            last = hs[-1]
            m = last.primitive[last.batch_index].material if last.batch_index >= 0 else last.primitive.material
            if m.opacity >= 1:
                opq = last
        origins.append(o); dirs.append(list(d))
        ids.append(rb.flat_prim_id(sc, item_ref[id(opq.primitive)], opq.batch_index) if opq else -1)
        dists.append(opq.dist if opq else 0)
        ntr.append(len(hs) - (1 if opq else 0))
    save('mixed3', sc, size=np.array([w, h]), float=ref_float(scene, w, h), packed=ref_packed(scene, w, h),
         ray_origins=np.array(origins, np.float32), ray_dirs=np.array(dirs, np.float32), ray_ids=np.array(ids, np.int32),
         ray_dists=np.array(dists, np.float32), ray_ntrans=np.array(ntr, np.int32))


def make_soup9():
    # 9-D simplex soup: above the reference's fixed-dimension modules, so it runs through the generic
    # `tracern` (var_geometry) path, and through the run-time-dimension kernels here (config 5 in miniature)
    dim = 9
    nt = ntracer.NTracer(dim)
    assert nt.base.__name__.endswith('tracern')
    rng = random.Random(1234)
    mat = ntracer.Material((1, 0.5, 0.5))
    mat2 = ntracer.Material((0.4, 0.8, 1.0), 1, 0.25)
    protos = []
    for i in range(300):
        c = [rng.uniform(-1, 1) for _ in range(3)] + [rng.uniform(-0.02, 0.02) for _ in range(dim - 3)]
        pts = [nt.Vector(*[ci + rng.uniform(-0.25, 0.25) for ci in c]) for _ in range(dim)]
        protos.append(nt.TrianglePrototype(pts, mat if i % 2 else mat2))
    scene = nt.build_composite_scene(protos)
    cam = nt.Camera()
    cam.translate(nt.Vector.axis(2, -3))
    scene.set_camera(cam)
    scene.add_light(nt.GlobalLight(nt.Vector.axis(1, -1) + nt.Vector.axis(2, 0.5), (0.6, 0.6, 0.6)))
    scene.set_shadows(True)
    scene.set_max_reflect_depth(2)
    w, h = 96, 54
    sc = rb.export_scene(nt, scene)
    ids, dist = ref_primary_ids(nt, scene, sc, w, h)
    save('soup9', sc, size=np.array([w, h]), float=ref_float(scene, w, h), ids=ids, dist=dist)


def make_rotation():
    """The interactive loop: frames of the reference itself along the camera path of
    scripts/polytope.py:522-557 (RotatingCamera), computed with the reference's own Camera / Matrix classes, for the
    exported cell120 (single-pass) and solids6 (reflective, wavefront passes) scenes.  40 steps per turn, 160x90."""
    import math
    from tests import fixtures as fx
    frames, w, h = 40, 160, 90
    out = {'frames': np.array(frames), 'size': np.array([w, h]), 'steps': np.array([5, 17, 29])}
    for name in ('cell120', 'solids6'):
        sc, g = fx.load(name)
        nt, scene, prims = rb.import_scene(sc)
        rb.make_immortal(prims.values())
        d = int(sc['dim'])
        cam = scene.get_camera()
        cam_distance = nt.dot(cam.origin, cam.axes[2])
        hh, incr = 1 / math.sqrt(d - 1), 2 * math.pi / frames
        cams_o, cams_a, packed = [], [], []
        for k in range(1, frames):
            a2 = cam.axes[0] * hh + cam.axes[1] * hh
            for i in range(d - 3):
                a2 += cam.axes[i + 3] * hh
            cam.transform(nt.Matrix.rotation(cam.axes[2], a2, incr))
            cam.normalize()
            cam.origin = cam.axes[2] * cam_distance
            if k in (5, 17, 29):
                scene.set_camera(cam)
                cams_o.append(np.array(list(cam.origin), np.float32))
                cams_a.append(np.array([list(cam.axes[i]) for i in range(d)], np.float32))
                packed.append(ref_packed(scene, w, h))
        out[name + '_origin'], out[name + '_axes'], out[name + '_packed'] = np.stack(cams_o), np.stack(cams_a), np.stack(packed)
    path = os.path.join(HERE, 'rotation.npz')
    np.savez_compressed(path, **out)
    print('wrote %s (%.1f KiB)' % (path, os.path.getsize(path) / 1024))


ALL = {'box': make_box, 'pack': make_pack, 'kdtree_kat': make_kdtree_kat, 'cell120': make_cell120,
       'ggs120': make_ggs120, 'ssc120': make_ssc120, 'solids6': make_solids6, 'mixed3': make_mixed3, 'soup9': make_soup9, 'rotation': make_rotation}

if __name__ == '__main__':
    random.seed(1)
    for name in (sys.argv[1:] or list(ALL)):
        print('==', name)
        ALL[name]()
