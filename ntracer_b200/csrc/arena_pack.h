// arena_pack.h -- host-side packing of an ntr_scene_desc into the single device arena
// (nodes | leaf items | simplex records | batch blocks | solid records | materials; every section 256-byte
// aligned, every record 16-byte aligned -- DESIGN.md section 3).
//
// Batch block (one per distinct triangle_batch item, B = batch_size lanes, B a multiple of 4):
//   [ face_normal[c][lane] : D*B ][ d[lane] : B ]            "plane part", SoA across lanes: one float4 = one
//                                                             component of 4 lanes, so the plane test of a lane group
//                                                             is D+1 coalesced 16-byte loads
//   [ lane 0: p1[D], edge_normals[D-1][D], pad to 4 ] ... [ lane B-1 ]   "edge parts", AoS: only read for lanes that
//                                                             survive the plane test
//   [ meta[lane] : B ]                                        material | opaque<<31  Header-only so that the test-only host
// emulation harness (tests/host_emul) packs scenes exactly like the product does.
#pragma once
#include <math.h>
#include <string.h>

#include <algorithm>

#include <unordered_map>
#include <vector>

#include "device_types.h"

namespace ntr {

struct ArenaLayout {
    size_t off_nodes = 0, off_refs = 0, off_simplex = 0, off_batches = 0, off_solids = 0, off_mats = 0, off_index = 0, total = 0;
    int sstride = 0, solstride = 0, lane_part = 0, batch_block = 0;
    bool any_transparent = false, any_reflective = false;
};

inline size_t arena_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline void pack_arena(const ntr_scene_desc *d, std::vector<unsigned char> &h, ArenaLayout &L) {
    const int D = d->dim;
    const int sin = (D + 1) * D + 1;
    L.sstride = (sin + 3) / 4 * 4;          // (D+1)*D+1 is odd: there is always a pad slot for the meta word
    const int solin = 1 + 2 * D * D + D;
    L.solstride = (solin + 1 + 3) / 4 * 4;
    const int B = d->batch_size;
    L.lane_part = (D * D + 3) / 4 * 4;
    L.batch_block = (D + 1) * B + L.lane_part * B + B;          // floats; a multiple of 4 because B is
    // distinct batch items (identity = leaf ref value) get one block each
    std::vector<uint32_t> batch_first;                          // first simplex index of every distinct batch
    std::unordered_map<uint32_t, uint32_t> batch_slot;
    for (uint32_t i = 0; i < d->n_leaf_refs; ++i) {
        const uint32_t r = d->leaf_refs[i];
        if ((r >> 30) == NTR_REF_BATCH && batch_slot.emplace(r, (uint32_t)batch_first.size()).second)
            batch_first.push_back(r & NTR_IDX_MASK);
    }
    L.off_nodes = 0;
    L.off_refs = arena_align(L.off_nodes + (size_t)d->n_nodes * 16, 256);
    L.off_simplex = arena_align(L.off_refs + (size_t)d->n_leaf_refs * 8, 256);
    L.off_batches = arena_align(L.off_simplex + (size_t)d->n_simplex * L.sstride * 4, 256);
    L.off_solids = arena_align(L.off_batches + batch_first.size() * (size_t)L.batch_block * 4, 256);
    L.off_mats = arena_align(L.off_solids + (size_t)d->n_solids * L.solstride * 4, 256);
    L.off_index = arena_align(L.off_mats + (size_t)d->n_materials * 12 * 4, 256);
    L.total = L.off_index;
    h.assign(L.total, 0);
    if (d->n_nodes) memcpy(h.data() + L.off_nodes, d->nodes, (size_t)d->n_nodes * 16);
    {   // leaf items: {ref, float offset of the record inside its section}
        uint32_t *it = reinterpret_cast<uint32_t *>(h.data() + L.off_refs);
        for (uint32_t i = 0; i < d->n_leaf_refs; ++i) {
            const uint32_t r = d->leaf_refs[i], kind = r >> 30, idx = r & NTR_IDX_MASK;
            uint32_t off;
            if (kind == NTR_REF_BATCH) off = batch_slot[r] * (uint32_t)L.batch_block;
            else if (kind == NTR_REF_SIMPLEX) off = idx * (uint32_t)L.sstride;
            else off = idx * (uint32_t)L.solstride;
            it[2 * i] = r;
            it[2 * i + 1] = off;
        }
    }
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
    float *bb = reinterpret_cast<float *>(h.data() + L.off_batches);
    for (size_t k = 0; k < batch_first.size(); ++k) {
        float *blk = bb + k * (size_t)L.batch_block;
        float *edge = blk + (D + 1) * B, *meta = edge + (size_t)L.lane_part * B;
        for (int l = 0; l < B; ++l) {
            const float *src = d->simplex + (size_t)(batch_first[k] + l) * sin;     // fn[D], d, p1[D], edges
            for (int c = 0; c <= D; ++c) blk[c * B + l] = src[c];
            memcpy(edge + (size_t)l * L.lane_part, src + D + 1, sizeof(float) * (size_t)D * D);
            const uint32_t m = mat_meta(d->simplex_mat[batch_first[k] + l]);
            memcpy(meta + l, &m, 4);
        }
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

// Mailbox keys (trace_core.cuh: MailboxStore) are record indices; a tree whose simplex items are all aligned 4-lane
// batches (what the reference's SSE builds produce) uses one key in four.  Returns 2 when the record indices of the
// leaf items stay distinct after dropping two bits -- the table then needs a quarter of the words -- else 0.
inline uint32_t mailbox_key_shift(const ntr_scene_desc *d) {
    std::unordered_map<uint32_t, uint32_t> owner;          // index >> 2  ->  the one leaf ref allowed to map there
    for (uint32_t i = 0; i < d->n_leaf_refs; ++i) {
        const uint32_t r = d->leaf_refs[i];
        if ((r >> 30) == NTR_REF_SOLID) continue;
        const auto it = owner.emplace((r & NTR_IDX_MASK) >> 2, r);
        if (!it.second && it.first->second != r) return 0;
    }
    return 2;
}

// Points a SceneDev at an arena that lives at `base` (device or, in the test harness, host memory).
inline void bind_arena(SceneDev &dev, const ntr_scene_desc *d, const ArenaLayout &L, const unsigned char *base) {
    dev.nodes = reinterpret_cast<const uint4 *>(base + L.off_nodes);
    dev.leaf_items = reinterpret_cast<const uint2 *>(base + L.off_refs);
    dev.batches = reinterpret_cast<const float *>(base + L.off_batches);
    dev.lane_part = L.lane_part;
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
