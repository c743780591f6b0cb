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
    int index_stride = 0;           // floats per leaf-index node: lo[D], hi[D], skip, item, padded to 4
    int sstride = 0, solstride = 0, lane_part = 0, batch_block = 0;
    bool any_transparent = false, any_reflective = false;
};

inline size_t arena_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Leaves with at least this many items get an in-order bounding-box index.  Measured (host emulation, 192x108):
// bit-identical images for opaque scenes, but on {5/2,3,3} it only removes 30 % of the simplex tests (the reference's
// leaf order is not spatially coherent and star-polytope cells have fat 4-D boxes) and the box tests cost more than
// they save, so it is OFF by default (build with -DNTR_LEAF_INDEX_MIN=12 -DNTR_USE_LEAF_INDEX=1 to experiment).
#ifndef NTR_LEAF_INDEX_MIN
#define NTR_LEAF_INDEX_MIN 0xFFFFFFFFu
#endif

// Axis-aligned bounds of the region one simplex record accepts: the vertices are p1 and p1 + v_j with
// edge_normal_i . v_j = -delta_ij and face_normal . v_j = 0 (checked against Triangle.from_points for D = 3..6);
// returns false for degenerate records (the caller then never culls them).
inline bool simplex_bounds(int D, const float *rec, double *lo, double *hi) {
    const float *fn = rec, *p1 = rec + D + 1, *edges = rec + 2 * D + 1;
    double M[NTR_MAXD][2 * NTR_MAXD];
    for (int r = 0; r < D; ++r) {
        const float *row = r < D - 1 ? edges + (size_t)r * D : fn;
        for (int c = 0; c < D; ++c) { M[r][c] = row[c]; M[r][D + c] = r == c ? 1.0 : 0.0; }
    }
    for (int c = 0; c < D; ++c) {                        // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < D; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (!(fabs(M[piv][c]) > 1e-30)) return false;
        if (piv != c) for (int k = 0; k < 2 * D; ++k) std::swap(M[piv][k], M[c][k]);
        const double inv = 1.0 / M[c][c];
        for (int k = 0; k < 2 * D; ++k) M[c][k] *= inv;
        for (int r = 0; r < D; ++r) {
            if (r == c) continue;
            const double f = M[r][c];
            if (f != 0) for (int k = 0; k < 2 * D; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int i = 0; i < D; ++i) { lo[i] = hi[i] = p1[i]; }
    for (int j = 0; j < D - 1; ++j) {
        for (int i = 0; i < D; ++i) {
            const double v = p1[i] - M[i][D + j];        // p1 + v_j, v_j = -(column j of M^-1)
            if (!(v == v) || fabs(v) > 1e30) return false;
            if (v < lo[i]) lo[i] = v;
            if (v > hi[i]) hi[i] = v;
        }
    }
    return true;
}

struct ItemBox { float lo[NTR_MAXD], hi[NTR_MAXD]; };

// Pre-order, in-order-visiting bounding-box index over the items of ONE leaf (DESIGN.md section 4 "leaf index"):
// node = {lo[D], hi[D], skip, item}; children follow their parent, `skip` is the index of the first node after the
// subtree, item >= 0 marks a single leaf item.  Walking it front to back and jumping to `skip` whenever the ray
// misses a box visits the surviving items in their ORIGINAL leaf order, which is what the reference's sequential
// leaf loop (and its first-tested-wins tie rule) needs.
inline void build_leaf_index(int D, int stride, const std::vector<ItemBox> &boxes, uint32_t a, uint32_t b,
                             std::vector<float> &out, uint32_t base) {
    const size_t me = out.size();
    out.resize(me + stride, 0.0f);
    if (b - a == 1) {
        for (int i = 0; i < D; ++i) { out[me + i] = boxes[a].lo[i]; out[me + D + i] = boxes[a].hi[i]; }
        const uint32_t skip = (uint32_t)((out.size() - base) / stride);
        memcpy(&out[me + 2 * D], &skip, 4);
        const int32_t item = (int32_t)a;
        memcpy(&out[me + 2 * D + 1], &item, 4);
        return;
    }
    const uint32_t mid = a + (b - a) / 2;
    build_leaf_index(D, stride, boxes, a, mid, out, base);
    build_leaf_index(D, stride, boxes, mid, b, out, base);
    for (int i = 0; i < D; ++i) {
        float lo = boxes[a].lo[i], hi = boxes[a].hi[i];
        for (uint32_t k = a + 1; k < b; ++k) { lo = std::min(lo, boxes[k].lo[i]); hi = std::max(hi, boxes[k].hi[i]); }
        out[me + i] = lo; out[me + D + i] = hi;
    }
    const uint32_t skip = (uint32_t)((out.size() - base) / stride);
    memcpy(&out[me + 2 * D], &skip, 4);
    const int32_t item = -1;
    memcpy(&out[me + 2 * D + 1], &item, 4);
}

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
    // leaf indices for big leaves
    L.index_stride = (2 * D + 2 + 3) / 4 * 4;
    std::vector<float> index_data;
    std::vector<uint32_t> node_index_off(d->n_nodes, 0);          // float4 offset + 1 into the index section, 0 = none
    {
        std::unordered_map<uint32_t, ItemBox> box_of;
        auto item_box = [&](uint32_t r) -> const ItemBox & {
            auto it = box_of.find(r);
            if (it != box_of.end()) return it->second;
            ItemBox bx;
            for (int i = 0; i < D; ++i) { bx.lo[i] = -3.0e38f; bx.hi[i] = 3.0e38f; }      // never culled
            const uint32_t kind = r >> 30, idx = r & NTR_IDX_MASK;
            if (kind != NTR_REF_SOLID) {
                const int lanes = kind == NTR_REF_BATCH ? B : 1;
                double lo[NTR_MAXD], hi[NTR_MAXD], tlo[NTR_MAXD], thi[NTR_MAXD];
                bool ok = true;
                for (int l = 0; l < lanes && ok; ++l) {
                    ok = simplex_bounds(D, d->simplex + (size_t)(idx + l) * sin, tlo, thi);
                    for (int i = 0; i < D && ok; ++i) {
                        if (l == 0) { lo[i] = tlo[i]; hi[i] = thi[i]; }
                        else { lo[i] = std::min(lo[i], tlo[i]); hi[i] = std::max(hi[i], thi[i]); }
                    }
                }
                if (ok) {
                    // the simplex test accepts points up to ROUNDING_FUZZ (barycentric) outside: pad generously
                    double scale = 1e-3;
                    for (int i = 0; i < D; ++i) scale = std::max(scale, std::max(fabs(lo[i]), fabs(hi[i])));
                    for (int i = 0; i < D; ++i) {
                        const double pad = 1e-4 * scale + 1e-4 * (hi[i] - lo[i]);
                        bx.lo[i] = (float)(lo[i] - pad);
                        bx.hi[i] = (float)(hi[i] + pad);
                    }
                }
            }
            return box_of.emplace(r, bx).first->second;
        };
        for (uint32_t n = 0; n < d->n_nodes; ++n) {
            const ntr_node &nd = d->nodes[n];
            if (!(nd.meta & NTR_LEAF_FLAG) || nd.w2 < NTR_LEAF_INDEX_MIN) continue;
            std::vector<ItemBox> boxes(nd.w2);
            for (uint32_t k = 0; k < nd.w2; ++k) boxes[k] = item_box(d->leaf_refs[nd.w1 + k]);
            const uint32_t base = (uint32_t)index_data.size();
            node_index_off[n] = base / 4 + 1;
            build_leaf_index(D, L.index_stride, boxes, 0, nd.w2, index_data, base);
        }
    }
    L.off_nodes = 0;
    L.off_refs = arena_align(L.off_nodes + (size_t)d->n_nodes * 16, 256);
    L.off_simplex = arena_align(L.off_refs + (size_t)d->n_leaf_refs * 8, 256);
    L.off_batches = arena_align(L.off_simplex + (size_t)d->n_simplex * L.sstride * 4, 256);
    L.off_solids = arena_align(L.off_batches + batch_first.size() * (size_t)L.batch_block * 4, 256);
    L.off_mats = arena_align(L.off_solids + (size_t)d->n_solids * L.solstride * 4, 256);
    L.off_index = arena_align(L.off_mats + (size_t)d->n_materials * 12 * 4, 256);
    L.total = arena_align(L.off_index + index_data.size() * 4, 256);
    h.assign(L.total, 0);
    if (d->n_nodes) memcpy(h.data() + L.off_nodes, d->nodes, (size_t)d->n_nodes * 16);
    for (uint32_t n = 0; n < d->n_nodes; ++n)           // leaf nodes: w3 = leaf index offset (0 = none)
        if (d->nodes[n].meta & NTR_LEAF_FLAG) memcpy(h.data() + L.off_nodes + (size_t)n * 16 + 12, &node_index_off[n], 4);
    if (!index_data.empty()) memcpy(h.data() + L.off_index, index_data.data(), index_data.size() * 4);
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

// Points a SceneDev at an arena that lives at `base` (device or, in the test harness, host memory).
inline void bind_arena(SceneDev &dev, const ntr_scene_desc *d, const ArenaLayout &L, const unsigned char *base) {
    dev.nodes = reinterpret_cast<const uint4 *>(base + L.off_nodes);
    dev.leaf_items = reinterpret_cast<const uint2 *>(base + L.off_refs);
    dev.batches = reinterpret_cast<const float *>(base + L.off_batches);
    dev.lane_part = L.lane_part;
    dev.leaf_index = reinterpret_cast<const float *>(base + L.off_index);
    dev.index_stride = L.index_stride;
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
