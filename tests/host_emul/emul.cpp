// TEST INFRASTRUCTURE ONLY -- runs the product's per-ray state machine (ntracer_b200/csrc/trace_core.cuh,
// the exact code the CUDA kernels inline) on the host, so that its logic can be checked against the
// oracle in the CPU-only test tier (`-m "not gpu"`) where no B200 exists.  It is compiled into
// tests/host_emul/libhostemul.so, which nothing under ntracer_b200/ loads: the product library
// libntracer_b200.so contains no host rendering code at all.
//
// The frame driver below mimics the kernel's pass structure (primary pass, then one pass per
// reflection depth over the queue written by the previous pass), single-threaded.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

// -DNTR_EMULATE_WARP (with -pthread): the warp-synchronous form of the per-ray path (trace_warp.cuh: cooperative leaf
// scans, shadow rays of a warp traced together; __shfl_sync / __ballot_sync / __any_sync / __reduce_*_sync between the
// 32 lanes of a warp) runs on the host too, each lane as an OS thread and every warp intrinsic as a rendezvous of the
// 32 threads.  Lanes of a correct kernel execute the
// same sequence of warp intrinsics, so a barrier per intrinsic reproduces the data exchange exactly; a divergent
// sequence shows up as a hang or a wrong image.
#ifdef NTR_EMULATE_WARP
#include <pthread.h>

#include <thread>
namespace warp_emu {
struct Warp {
    pthread_barrier_t bar;
    uint32_t slot[32];
    Warp() { pthread_barrier_init(&bar, nullptr, 32); }
    ~Warp() { pthread_barrier_destroy(&bar); }
};
thread_local int lane = 0;
thread_local Warp *warp = nullptr;
struct Idx { int x; };
inline Idx tidx() { return Idx{lane}; }
}  // namespace warp_emu
#define threadIdx (warp_emu::tidx())
template <typename T> inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    memcpy(&warp_emu::warp->slot[warp_emu::lane], &v, 4);
    pthread_barrier_wait(&warp_emu::warp->bar);
    T r;
    memcpy(&r, &warp_emu::warp->slot[src & 31], 4);
    pthread_barrier_wait(&warp_emu::warp->bar);
    return r;
}
template <typename T> inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) {
    return __shfl_sync(m, v, warp_emu::lane ^ lane_mask);
}
inline unsigned __ballot_sync(unsigned, bool p) {
    warp_emu::warp->slot[warp_emu::lane] = p ? 1u : 0u;
    pthread_barrier_wait(&warp_emu::warp->bar);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= warp_emu::warp->slot[i] << i;
    pthread_barrier_wait(&warp_emu::warp->bar);
    return m;
}
inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) == 0xFFFFFFFFu; }
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
template <typename F> inline unsigned warp_reduce_u32(unsigned v, F f) {
    warp_emu::warp->slot[warp_emu::lane] = v;
    pthread_barrier_wait(&warp_emu::warp->bar);
    unsigned r = warp_emu::warp->slot[0];
    for (int i = 1; i < 32; ++i) r = f(r, warp_emu::warp->slot[i]);
    pthread_barrier_wait(&warp_emu::warp->bar);
    return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return warp_reduce_u32(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return warp_reduce_u32(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline void __syncwarp() { pthread_barrier_wait(&warp_emu::warp->bar); }
inline unsigned __match_any_sync(unsigned, unsigned v) {
    warp_emu::warp->slot[warp_emu::lane] = v;
    pthread_barrier_wait(&warp_emu::warp->bar);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (warp_emu::warp->slot[i] == v ? 1u : 0u) << i;
    pthread_barrier_wait(&warp_emu::warp->bar);
    return m;
}
inline int atomicAdd(int *p, int v) { return __sync_fetch_and_add(p, v); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return warp_reduce_u32(v, [](unsigned a, unsigned b) { return a + b; }); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
#define __device__
#define __forceinline__ inline
#include <vector_types.h>
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

#include "../../ntracer_b200/csrc/arena_pack.h"
#include "../../ntracer_b200/csrc/trace_core.cuh"
#ifdef NTR_EMULATE_WARP
#include "../../ntracer_b200/csrc/trace_warp.cuh"
#endif

using namespace ntr;

namespace {

template <int DT> struct VecEmit {
    std::vector<Bounce<DT>> *out;
    std::vector<uint32_t> *pix;
    uint32_t pixel;
    void operator()(const Bounce<DT> &b) { out->push_back(b); pix->push_back(pixel); }
};

struct Scene {
    std::vector<unsigned char> arena;
    std::vector<float> lights;
    std::vector<uint32_t> mailbox;      // the exact mailbox table (capi.cu allocates it under the same conditions)
    SceneDev dev;
    CameraDev cam;
    int flags;
};

void setup(Scene &S, const ntr_scene_desc *d, const float *cam_origin, const float *cam_axes) {
    memset(&S.dev, 0, sizeof S.dev);
    memset(&S.cam, 0, sizeof S.cam);
    fill_scene_params(S.dev, d, 63);
    S.dev.root = NTR_NULL_NODE;
    S.dev.batch = 1;
    S.flags = 0;
    if (d->kind == NTR_SCENE_COMPOSITE) {
        ArenaLayout L;
        pack_arena(d, S.arena, L);
        bind_arena(S.dev, d, L, S.arena.data());
        S.flags = (L.any_transparent || d->n_solids) ? NTR_F_GENERAL : 0;
        const size_t stride = d->dim + 3;
        S.lights.resize(((size_t)d->n_point_lights + d->n_global_lights) * stride + 1);
        if (d->n_point_lights) memcpy(S.lights.data(), d->point_lights, sizeof(float) * d->n_point_lights * stride);
        if (d->n_global_lights) memcpy(S.lights.data() + d->n_point_lights * stride, d->global_lights, sizeof(float) * d->n_global_lights * stride);
        S.dev.point_lights = S.lights.data();
        S.dev.global_lights = S.lights.data() + d->n_point_lights * stride;
        S.dev.n_point = (int)d->n_point_lights;
        S.dev.n_global = (int)d->n_global_lights;
        uint32_t max_leaf = 0;
        for (uint32_t i = 0; i < d->n_nodes; ++i) if (d->nodes[i].meta & NTR_LEAF_FLAG) max_leaf = std::max(max_leaf, d->nodes[i].w2);
        const uint64_t keys = (uint64_t)d->n_simplex + d->n_solids;
#ifndef NTR_EMUL_EXACT_MAILBOX_ALWAYS
#define NTR_EMUL_EXACT_MAILBOX_ALWAYS 0
#endif
        if ((S.flags & NTR_F_GENERAL) && (max_leaf > NTR_MAILBOX_CAP || NTR_EMUL_EXACT_MAILBOX_ALWAYS) && keys <= NTR_MAILBOX_MAX_KEYS) {
            S.dev.mb_threads = 32;
            S.dev.mb_shift = mailbox_key_shift(d);
            const uint64_t nk = ((uint64_t)d->n_simplex >> S.dev.mb_shift) + 1 + d->n_solids;
            S.dev.mb_words = (uint32_t)((nk + NTR_MAILBOX_BITS_PER_WORD - 1) / NTR_MAILBOX_BITS_PER_WORD);
            S.mailbox.assign((size_t)(S.dev.mb_words + 1) * S.dev.mb_threads, 0u);
            S.dev.mb_table = S.mailbox.data();
        }
    }
    if (cam_origin && cam_axes) {
        for (int i = 0; i < d->dim; ++i) {
            S.cam.origin[i] = cam_origin[i];
            S.cam.right[i] = cam_axes[i]; S.cam.up[i] = cam_axes[d->dim + i]; S.cam.fwd[i] = cam_axes[2 * d->dim + i];
        }
    }
}

template <int DT, int FLAGS>
void render_t(Scene &S, int w, int h, float *rgb, int32_t *ids, float *dists, unsigned long long *cnt_out) {
    constexpr int CAP = DimCap<DT>::value;
    FrameDev f;
    memset(&f, 0, sizeof f);
    f.width = w; f.height = h;
    f.half_w = (float)w / 2.0f; f.half_h = (float)h / 2.0f;
    f.fovI = tanf(S.dev.fov / 2) / f.half_w;
    Counters cnt;
    std::vector<Bounce<DT>> q, qn;
    std::vector<uint32_t> qp, qpn;
    const float one[3] = {1, 1, 1};
    MailboxStore ms;
    ms.attach(S.dev.mb_table, S.dev.mb_words, S.dev.mb_threads, 0, S.dev.n_simplex, S.dev.mb_shift);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const uint32_t pix = (uint32_t)y * w + x;
            float o[CAP], dir[CAP], acc[3] = {0, 0, 0};
            HitRec prim;
            prim.dist = 0; prim.ref = NTR_NONE_REF; prim.lane = -1;
            primary_ray<DT>(S.dev, S.cam, f, x, y, o, dir);
            if (S.dev.kind == NTR_SCENE_BOX) box_color<DT>(S.dev, o, dir, acc, &prim);
            else {
                VecEmit<DT> emit{&q, &qp, pix};
                const Skip none = {NTR_NONE_REF, 0};
                ray_color<DT, FLAGS>(S.dev, true, o, dir, 0, none, one, acc, emit, cnt, &prim, &ms);
            }
            if (rgb) { rgb[pix * 3] = acc[0]; rgb[pix * 3 + 1] = acc[1]; rgb[pix * 3 + 2] = acc[2]; }
            if (ids) ids[pix] = prim.ref == NTR_NONE_REF ? -1 : (S.dev.kind == NTR_SCENE_BOX ? 0 : flat_prim_id(S.dev, prim.ref, prim.lane));
            if (dists) dists[pix] = prim.dist;
        }
    }
    while (!q.empty() && rgb) {
        qn.clear(); qpn.clear();
        for (size_t i = 0; i < q.size(); ++i) {
            float acc[3] = {0, 0, 0};
            VecEmit<DT> emit{&qn, &qpn, qp[i]};
            ray_color<DT, FLAGS>(S.dev, true, q[i].o, q[i].d, q[i].depth, q[i].skip, q[i].w, acc, emit, cnt, nullptr, &ms);
            rgb[(size_t)qp[i] * 3] += acc[0]; rgb[(size_t)qp[i] * 3 + 1] += acc[1]; rgb[(size_t)qp[i] * 3 + 2] += acc[2];
        }
        q.swap(qn); qp.swap(qpn);
    }
    ms.detach();
    if (cnt_out) {
        cnt_out[0] = (unsigned long long)w * h; cnt_out[1] = cnt.reflection_rays; cnt_out[2] = cnt.shadow_rays;
        cnt_out[3] = cnt.node_steps; cnt_out[4] = cnt.simplex_tests; cnt_out[5] = cnt.solid_tests; cnt_out[6] = cnt.shaded_hits;
        cnt_out[7] = 0;
        cnt_out[8] = cnt.truncated;
    }
}

#ifdef NTR_EMULATE_WARP
// The same frame, 32 rays at a time: every group of 32 consecutive rays is a warp whose lanes run as threads.
template <int DT> struct LaneEmit {
    std::vector<std::pair<size_t, Bounce<DT>>> *out;
    std::vector<uint32_t> *pix;
    size_t order;
    uint32_t pixel;
    void operator()(const Bounce<DT> &b) { out->push_back({order, b}); pix->push_back(pixel); }
};

template <int DT, int FLAGS>
void render_warp_t(Scene &S, int w, int h, float *rgb, int32_t *ids, float *dists, unsigned long long *cnt_out) {
    constexpr int CAP = DimCap<DT>::value;
    FrameDev f;
    memset(&f, 0, sizeof f);
    f.width = w; f.height = h;
    f.half_w = (float)w / 2.0f; f.half_h = (float)h / 2.0f;
    f.fovI = tanf(S.dev.fov / 2) / f.half_w;
    const float one[3] = {1, 1, 1};
    std::vector<Bounce<DT>> q;          // rays of the current pass (empty = primary pass)
    std::vector<uint32_t> qp;
    Counters total;
    bool primary = true;
    while (primary || !q.empty()) {
        const size_t n = primary ? (size_t)w * h : q.size();
        warp_emu::Warp W;
        int done_ctr = 0;               // the per-warp counter the kernels keep in shared memory
        std::vector<std::pair<size_t, Bounce<DT>>> out[32];
        std::vector<uint32_t> outpix[32];
        Counters cnts[32];
        std::vector<std::thread> lanes;
        for (int L = 0; L < 32; ++L) {
            lanes.emplace_back([&, L]() {
                warp_emu::lane = L;
                warp_emu::warp = &W;
                MailboxStore ms;
                ms.attach(S.dev.mb_table, S.dev.mb_words, S.dev.mb_threads, (uint32_t)L, S.dev.n_simplex, S.dev.mb_shift);
                for (size_t base = 0; base < n; base += 32) {
                    const size_t idx = base + L;
                    const bool enabled = idx < n;
                    float o[CAP], dir[CAP], acc[3] = {0, 0, 0}, wgt[3] = {1, 1, 1};
                    for (int k = 0; k < CAP; ++k) { o[k] = 0; dir[k] = 1; }
                    Skip skip = {NTR_NONE_REF, 0};
                    int depth = 0;
                    uint32_t pix = 0;
                    HitRec prim;
                    prim.dist = 0; prim.ref = NTR_NONE_REF; prim.lane = -1;
                    if (enabled) {
                        if (primary) {
                            pix = (uint32_t)idx;
                            primary_ray<DT>(S.dev, S.cam, f, (int)(idx % w), (int)(idx / w), o, dir);
                        } else {
                            pix = qp[idx];
                            for (int k = 0; k < CAP; ++k) { o[k] = q[idx].o[k]; dir[k] = q[idx].d[k]; }
                            depth = q[idx].depth; skip = q[idx].skip;
                            wgt[0] = q[idx].w[0]; wgt[1] = q[idx].w[1]; wgt[2] = q[idx].w[2];
                        }
                    }
                    LaneEmit<DT> emit{&out[L], &outpix[L], idx, pix};
                    ray_color_warp<DT, FLAGS>(S.dev, enabled, o, dir, depth, skip, primary ? one : wgt, acc, emit, cnts[L], primary ? &prim : nullptr, &done_ctr, &ms);
                    if (!enabled) continue;
                    if (primary) {
                        if (rgb) { rgb[(size_t)pix * 3] = acc[0]; rgb[(size_t)pix * 3 + 1] = acc[1]; rgb[(size_t)pix * 3 + 2] = acc[2]; }
                        if (ids) ids[pix] = prim.ref == NTR_NONE_REF ? -1 : flat_prim_id(S.dev, prim.ref, prim.lane);
                        if (dists) dists[pix] = prim.dist;
                    }
                    // (bounce results are added below, in ray order, so that the float sums match the sequential driver)
                    else { out[L].push_back({idx, Bounce<DT>{}}); outpix[L].push_back(0xFFFFFFFFu); out[L].back().second.w[0] = acc[0]; out[L].back().second.w[1] = acc[1]; out[L].back().second.w[2] = acc[2]; }
                }
                ms.detach();
            });
        }
        for (auto &t : lanes) t.join();
        // merge the lanes' emissions in ray order (= the order the sequential driver produces)
        struct Item { size_t order; int lane; size_t k; };
        std::vector<Item> items;
        for (int L = 0; L < 32; ++L)
            for (size_t k = 0; k < out[L].size(); ++k) items.push_back({out[L][k].first, L, k});
        std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.order < b.order; });
        std::vector<Bounce<DT>> qn;
        std::vector<uint32_t> qpn;
        for (const Item &it : items) {
            const uint32_t px = outpix[it.lane][it.k];
            const Bounce<DT> &b = out[it.lane][it.k].second;
            if (px == 0xFFFFFFFFu) {                 // result of bounce ray `order`
                if (rgb) { const uint32_t p = qp[it.order]; rgb[(size_t)p * 3] += b.w[0]; rgb[(size_t)p * 3 + 1] += b.w[1]; rgb[(size_t)p * 3 + 2] += b.w[2]; }
            } else { qn.push_back(b); qpn.push_back(px); }
        }
        for (int L = 0; L < 32; ++L) {
            total.reflection_rays += cnts[L].reflection_rays; total.shadow_rays += cnts[L].shadow_rays;
            total.node_steps += cnts[L].node_steps; total.simplex_tests += cnts[L].simplex_tests;
            total.solid_tests += cnts[L].solid_tests; total.shaded_hits += cnts[L].shaded_hits; total.truncated += cnts[L].truncated;
        }
        q.swap(qn); qp.swap(qpn);
        primary = false;
        if (!rgb) break;
    }
    if (cnt_out) {
        cnt_out[0] = (unsigned long long)w * h; cnt_out[1] = total.reflection_rays; cnt_out[2] = total.shadow_rays;
        cnt_out[3] = total.node_steps; cnt_out[4] = total.simplex_tests; cnt_out[5] = total.solid_tests; cnt_out[6] = total.shaded_hits;
        cnt_out[7] = 0;
        cnt_out[8] = total.truncated;
    }
}
#endif

template <int DT> void render_d(Scene &S, int w, int h, float *rgb, int32_t *ids, float *dists, unsigned long long *cnt) {
#ifdef NTR_EMULATE_WARP
    if (S.dev.kind != NTR_SCENE_BOX) {
        if (S.flags & NTR_F_GENERAL) render_warp_t<DT, NTR_F_GENERAL | NTR_F_COUNT>(S, w, h, rgb, ids, dists, cnt);
        else render_warp_t<DT, NTR_F_COUNT>(S, w, h, rgb, ids, dists, cnt);
        return;
    }
#endif
    if (S.flags & NTR_F_GENERAL) render_t<DT, NTR_F_GENERAL | NTR_F_COUNT>(S, w, h, rgb, ids, dists, cnt);
    else render_t<DT, NTR_F_COUNT>(S, w, h, rgb, ids, dists, cnt);
}

template <int DT, int FLAGS>
void trace_t(Scene &S, uint32_t n, const float *origins, const float *dirs, float t_near, float t_far,
             const uint32_t *skip_ref, const int32_t *skip_lane, int32_t *ids, float *dist, int32_t *ntrans,
             int max_hits = 0, int32_t *hit_ids = nullptr, float *hit_dists = nullptr) {
    const int D = S.dev.dim;
    MailboxStore ms;
    ms.attach((FLAGS & NTR_F_GENERAL) ? S.dev.mb_table : nullptr, S.dev.mb_words, S.dev.mb_threads, 0, S.dev.n_simplex, S.dev.mb_shift);
    for (uint32_t i = 0; i < n; ++i) {
        Skip skip = {skip_ref ? skip_ref[i] : NTR_NONE_REF, skip_lane ? skip_lane[i] : -1};
        GenState<DT> g;
        g.mb.big = ms.col ? &ms : nullptr;
        g.th.clear();
        HitRec oh;
        oh.dist = FLT_MAX; oh.ref = NTR_NONE_REF; oh.lane = -1;
        Counters cnt;
        const bool hit = trace_nearest<DT, FLAGS>(S.dev, origins + (size_t)i * D, dirs + (size_t)i * D, skip, t_near, t_far, oh, &g, cnt);
        ids[i] = hit ? flat_prim_id(S.dev, oh.ref, oh.lane) : -1;
        if (dist) dist[i] = hit ? oh.dist : 0;
        if (ntrans) ntrans[i] = (FLAGS & NTR_F_GENERAL) ? g.th.n : 0;
        if ((FLAGS & NTR_F_GENERAL) && hit_ids) {
            for (int k = 0; k < g.th.n && k < max_hits; ++k) {
                hit_ids[(size_t)i * max_hits + k] = flat_prim_id(S.dev, g.th.ref[k], g.th.lane[k]);
                if (hit_dists) hit_dists[(size_t)i * max_hits + k] = g.th.dist[k];
            }
        }
    }
}

template <int DT, int FLAGS>
void occl_t(Scene &S, uint32_t n, const float *origins, const float *dirs, const float *distance,
            const uint32_t *skip_ref, const int32_t *skip_lane, int32_t *occ, int32_t *ntrans) {
    const int D = S.dev.dim;
    for (uint32_t i = 0; i < n; ++i) {
        Skip skip = {skip_ref ? skip_ref[i] : NTR_NONE_REF, skip_lane ? skip_lane[i] : -1};
        HitList hits;
        hits.clear();
        Counters cnt;
        const bool r = trace_occludes<DT, FLAGS>(S.dev, origins + (size_t)i * D, dirs + (size_t)i * D,
                                                distance ? distance[i] : FLT_MAX, skip, -FLT_MAX, FLT_MAX, &hits, cnt);
        occ[i] = r;
        if (ntrans) ntrans[i] = (!r && (FLAGS & NTR_F_GENERAL)) ? hits.n : 0;
    }
}

#define DISPATCH_DIM(dim, CALL)            \
    switch (dim) {                         \
        case 3: { CALL(3); break; }        \
        case 4: { CALL(4); break; }        \
        case 5: { CALL(5); break; }        \
        case 6: { CALL(6); break; }        \
        case 7: { CALL(7); break; }        \
        case 8: { CALL(8); break; }        \
        case 9: { CALL(9); break; }        \
        case 10: { CALL(10); break; }      \
        default: { CALL(0); break; }       \
    }

}  // namespace

extern "C" {

// force_generic != 0 runs the run-time-dimension instantiation (DT = 0) whatever the dimension is
__attribute__((visibility("default"))) int emul_render(const ntr_scene_desc *d, const float *cam_origin,
                                                        const float *cam_axes, int w, int h, int force_generic,
                                                        float *rgb, int32_t *ids, float *dists,
                                                        unsigned long long *counters) {
    Scene S;
    setup(S, d, cam_origin, cam_axes);
#define CALL(DT) render_d<DT>(S, w, h, rgb, ids, dists, counters)
    DISPATCH_DIM(force_generic ? 0 : d->dim, CALL)
#undef CALL
    return 0;
}

__attribute__((visibility("default"))) int emul_trace_rays(const ntr_scene_desc *d, uint32_t n, const float *origins,
                                                            const float *dirs, float t_near, float t_far,
                                                            const uint32_t *skip_ref, const int32_t *skip_lane,
                                                            int force_generic, int32_t *ids, float *dist, int32_t *ntrans) {
    Scene S;
    setup(S, d, nullptr, nullptr);
#define CALL(DT)                                                                                                  \
    if (S.flags & NTR_F_GENERAL) trace_t<DT, NTR_F_GENERAL>(S, n, origins, dirs, t_near, t_far, skip_ref, skip_lane, ids, dist, ntrans); \
    else trace_t<DT, 0>(S, n, origins, dirs, t_near, t_far, skip_ref, skip_lane, ids, dist, ntrans)
    DISPATCH_DIM(force_generic ? 0 : d->dim, CALL)
#undef CALL
    return 0;
}

__attribute__((visibility("default"))) int emul_trace_rays_hits(const ntr_scene_desc *d, uint32_t n, const float *origins,
                                                                 const float *dirs, float t_near, float t_far,
                                                                 const uint32_t *skip_ref, const int32_t *skip_lane,
                                                                 int force_generic, int32_t *ids, float *dist, int32_t *ntrans,
                                                                 int max_hits, int32_t *hit_ids, float *hit_dists) {
    Scene S;
    setup(S, d, nullptr, nullptr);
    for (size_t k = 0; k < (size_t)n * max_hits; ++k) hit_ids[k] = -1;
#define CALL(DT)                                                                                                  \
    if (S.flags & NTR_F_GENERAL) trace_t<DT, NTR_F_GENERAL>(S, n, origins, dirs, t_near, t_far, skip_ref, skip_lane, ids, dist, ntrans, max_hits, hit_ids, hit_dists); \
    else trace_t<DT, 0>(S, n, origins, dirs, t_near, t_far, skip_ref, skip_lane, ids, dist, ntrans)
    DISPATCH_DIM(force_generic ? 0 : d->dim, CALL)
#undef CALL
    return 0;
}

__attribute__((visibility("default"))) int emul_occludes_rays(const ntr_scene_desc *d, uint32_t n, const float *origins,
                                                               const float *dirs, const float *distance,
                                                               const uint32_t *skip_ref, const int32_t *skip_lane,
                                                               int force_generic, int32_t *occ, int32_t *ntrans) {
    Scene S;
    setup(S, d, nullptr, nullptr);
#define CALL(DT)                                                                                              \
    if (S.flags & NTR_F_GENERAL) occl_t<DT, NTR_F_GENERAL>(S, n, origins, dirs, distance, skip_ref, skip_lane, occ, ntrans); \
    else occl_t<DT, 0>(S, n, origins, dirs, distance, skip_ref, skip_lane, occ, ntrans)
    DISPATCH_DIM(force_generic ? 0 : d->dim, CALL)
#undef CALL
    return 0;
}

__attribute__((visibility("default"))) int emul_pack(const ntr_image_format *fmt, const float *rgb, unsigned char *dst) {
    FormatDev f;
    memset(&f, 0, sizeof f);
    f.n_channels = fmt->n_channels; f.bytes_per_pixel = fmt->bytes_per_pixel; f.reversed = fmt->reversed; f.pitch = fmt->pitch;
    for (int i = 0; i < fmt->n_channels; ++i) {
        f.f_r[i] = fmt->channels[i].f_r; f.f_g[i] = fmt->channels[i].f_g; f.f_b[i] = fmt->channels[i].f_b; f.f_c[i] = fmt->channels[i].f_c;
        f.bits[i] = fmt->channels[i].bit_size; f.tfloat[i] = fmt->channels[i].tfloat;
    }
    const int bpp = f.bytes_per_pixel;
    for (int y = 0; y < fmt->height; ++y)
        for (int x = 0; x < fmt->width; ++x) {
            uint32_t w[4];
            pack_pixel(f, rgb + ((size_t)y * fmt->width + x) * 3, w);
            unsigned char *p = dst + (size_t)y * fmt->pitch + (size_t)x * bpp;
            for (int j = 0; j < bpp; ++j) {
                const int sj = f.reversed ? bpp - 1 - j : j;
                p[j] = (unsigned char)(w[sj >> 2] >> (8 * (3 - (sj & 3))));
            }
        }
    return 0;
}

}  // extern "C"
