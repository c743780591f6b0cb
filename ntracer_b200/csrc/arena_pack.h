// arena_pack.h -- host-side packing of an ntr_scene_desc into the single device arena
// (nodes | leaf refs | simplex records | solid records | materials; every section 256-byte aligned,
// every record 16-byte aligned -- DESIGN.md section 3).  Header-only so that the test-only host
// emulation harness (tests/host_emul) packs scenes exactly like the product does.
#pragma once
#include <string.h>

#include <vector>

#include "device_types.h"

namespace ntr {

struct ArenaLayout {
    size_t off_nodes = 0, off_refs = 0, off_simplex = 0, off_solids = 0, off_mats = 0, total = 0;
    int sstride = 0, solstride = 0;
    bool any_transparent = false, any_reflective = false;
};

inline size_t arena_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline void pack_arena(const ntr_scene_desc *d, std::vector<unsigned char> &h, ArenaLayout &L) {
    const int D = d->dim;
    const int sin = (D + 1) * D + 1;
    L.sstride = (sin + 3) / 4 * 4;          // (D+1)*D+1 is odd: there is always a pad slot for the meta word
    const int solin = 1 + 2 * D * D + D;
    L.solstride = (solin + 1 + 3) / 4 * 4;
    L.off_nodes = 0;
    L.off_refs = arena_align(L.off_nodes + (size_t)d->n_nodes * 16, 256);
    L.off_simplex = arena_align(L.off_refs + (size_t)d->n_leaf_refs * 4, 256);
    L.off_solids = arena_align(L.off_simplex + (size_t)d->n_simplex * L.sstride * 4, 256);
    L.off_mats = arena_align(L.off_solids + (size_t)d->n_solids * L.solstride * 4, 256);
    L.total = arena_align(L.off_mats + (size_t)d->n_materials * 12 * 4, 256);
    h.assign(L.total, 0);
    if (d->n_nodes) memcpy(h.data() + L.off_nodes, d->nodes, (size_t)d->n_nodes * 16);
    if (d->n_leaf_refs) memcpy(h.data() + L.off_refs, d->leaf_refs, (size_t)d->n_leaf_refs * 4);
    auto mat_meta = [&](int32_t m) -> uint32_t {
        const float *mm = d->materials + (size_t)m * 10;
        const bool opaque = mm[6] >= 1.0f;              // primitive::opaque, reference src/tracer.hpp:187-189
        if (!opaque) L.any_transparent = true;
        if (mm[7] != 0.0f) L.any_reflective = true;
        return (uint32_t)m | (opaque ? NTR_META_OPAQUE : 0u);
    };
    float *sx = reinterpret_cast<float *>(h.data() + L.off_simplex);
    for (uint32_t i = 0; i < d->n_simplex; ++i) {
        memcpy(sx + (size_t)i * L.sstride, d->simplex + (size_t)i * sin, sizeof(float) * sin);
        const uint32_t meta = mat_meta(d->simplex_mat[i]);
        memcpy(sx + (size_t)i * L.sstride + L.sstride - 1, &meta, 4);
    }
    float *so = reinterpret_cast<float *>(h.data() + L.off_solids);
    for (uint32_t i = 0; i < d->n_solids; ++i) {
        const float *src = d->solids + (size_t)i * solin;
        float *dst = so + (size_t)i * L.solstride;
        dst[0] = src[0];
        memcpy(dst + 1, src + 1 + D * D, sizeof(float) * D * D);                // inv_orientation
        memcpy(dst + 1 + D * D, src + 1 + 2 * D * D, sizeof(float) * D);        // position
        memcpy(dst + 1 + D * D + D, src + 1, sizeof(float) * D * D);            // orientation
        const uint32_t meta = mat_meta(d->solid_mat[i]);
        memcpy(dst + L.solstride - 1, &meta, 4);
    }
    float *mt = reinterpret_cast<float *>(h.data() + L.off_mats);
    for (uint32_t i = 0; i < d->n_materials; ++i)
        memcpy(mt + (size_t)i * 12, d->materials + (size_t)i * 10, sizeof(float) * 10);
}

// Points a SceneDev at an arena that lives at `base` (device or, in the test harness, host memory).
inline void bind_arena(SceneDev &dev, const ntr_scene_desc *d, const ArenaLayout &L, const unsigned char *base) {
    dev.nodes = reinterpret_cast<const uint4 *>(base + L.off_nodes);
    dev.leaf_refs = reinterpret_cast<const uint32_t *>(base + L.off_refs);
    dev.simplex = reinterpret_cast<const float *>(base + L.off_simplex);
    dev.solids = reinterpret_cast<const float *>(base + L.off_solids);
    dev.materials = reinterpret_cast<const float *>(base + L.off_mats);
    dev.sstride = L.sstride;
    dev.solstride = L.solstride;
    dev.root = d->root;
    dev.n_simplex = d->n_simplex;
    dev.batch = d->batch_size;
}

// composite_scene state (reference src/tracer.hpp:1713-1725) -> SceneDev
inline void fill_scene_params(SceneDev &dev, const ntr_scene_desc *d, int max_depth_cap) {
    dev.dim = d->dim;
    dev.kind = d->kind;
    dev.fov = d->fov;
    if (d->kind != NTR_SCENE_COMPOSITE) return;
    dev.shadows = d->shadows;
    dev.camera_light = d->camera_light;
    dev.max_depth = d->max_reflect_depth < max_depth_cap ? d->max_reflect_depth : max_depth_cap;
    dev.bg_axis = d->bg_gradient_axis;
    for (int c = 0; c < 3; ++c) {
        dev.ambient[c] = d->ambient[c]; dev.bg1[c] = d->bg1[c]; dev.bg2[c] = d->bg2[c]; dev.bg3[c] = d->bg3[c];
    }
    for (int i = 0; i < d->dim; ++i) { dev.bmin[i] = d->boundary[i]; dev.bmax[i] = d->boundary[d->dim + i]; }
}

}  // namespace ntr
