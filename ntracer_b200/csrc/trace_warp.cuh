// trace_warp.cuh -- the warp-synchronous form of the per-ray path: what the render kernels run.
//
// trace_core.cuh states the per-ray algorithm one ray at a time (that form is what the CPU tier checks against the
// oracle).  Here the 32 rays of a warp go through the same steps TOGETHER, which buys two things:
//
//  1. Warp-cooperative big leaves (north star (b): traversal with warp votes).  A ray that reaches a leaf with at least
//     NTR_COOP_LEAF_MIN items parks there.  When every lane of the warp is either finished or parked, one ballot
//     decides: if splitting the parked leaves over the 32 lanes (sum of ceil(size/32) chunks) is cheaper than every
//     parked lane scanning its own leaf (max size), the leaves are served one ray at a time -- the owner's ray is
//     broadcast, lane j tests item base+j of each 32-item chunk, and the owner folds the results back in leaf order;
//     otherwise (coherent rays parked at the same leaf) the lanes scan for themselves as before.  A single ray through
//     the 1,600-item leaves of {5/2,3,3} costs 1/32 of what it did whenever its warp-mates have nothing left to do,
//     which is what bounded the tail of every wavefront pass and of every frame spread over 8 GPUs.
//     Exactness: opaque variant -- the winner is the lexicographic minimum of (t, item index, batch lane), which is
//     what the sequential scan with its strict `t < cutoff` keeps; general variant -- chunk evaluation + in-order
//     replay (trace_core.cuh: replay_item), proven bit-identical to the item-by-item scan in the CPU tier; occlusion --
//     any opaque hit decides, transparent blockers are appended in leaf order.
//  2. Shadow rays of a warp are traced together (north star (d)): the layers of a pixel and the lights of a layer are
//     walked in warp-uniform loops, every lane that needs the shadow ray of (layer i, light l) enters the occlusion
//     traversal with the others (a ballot-compacted wavefront inside the warp, including cooperative big leaves),
//     and the lights of one hit are still applied in light order (append_specular is order dependent).
//
// Everything here must be called by all 32 lanes of a warp, converged; lanes without a ray pass enabled = false.
// tests/host_emul runs this file on an emulated warp (32 host threads, every warp intrinsic a rendezvous) and checks
// it bit for bit against the per-ray form.
#pragma once
#include "trace_core.cuh"

#if NTR_WARP_CODE
namespace ntr {

#ifndef NTR_COOP_LEAF_MIN
#define NTR_COOP_LEAF_MIN 48        // leaves with at least this many items park the ray for the warp's decision
#endif
#ifndef NTR_COOP_OVERHEAD
#define NTR_COOP_OVERHEAD 3         // cost of serving one parked ray (broadcast + fold), in units of one item test
#endif
#ifndef NTR_COOP_CHUNK_COST
#define NTR_COOP_CHUNK_COST 3       // cost of one cooperative 32-item chunk (test + votes + fold), in units of one item test
                                    // (measured, config 4: 1 -> 45.6 ms, 2 -> 44.0, 3 -> 43.5; the other thresholds move it by < 1 %)
#endif
#ifndef NTR_COOP_MIN_DONE
#define NTR_COOP_MIN_DONE 16        // a ray parks at a big leaf only while at least this many lanes of its warp have nothing
                                    // left to do (idle lanes are what cooperation feeds on; parking in a busy warp
                                    // stalls the parked lane until every other lane has finished or parked too)
#endif
#ifndef NTR_WARP_SHADE
#define NTR_WARP_SHADE 0            // 1: the layers and lights of the warp's rays are walked in warp-uniform loops and their
                                    // shadow rays traced together (cooperative big leaves included); 0: every lane shades its
                                    // own hits after the warp's nearest-hit traversal.  Measured (B200, round 2 call 4): 1 is
                                    // slower everywhere -- config 4 121 vs 109 ms, its 1/8 share 24.4 vs 22.9 ms, config 5
                                    // reduced 36.1 vs 30.3 ms -- so it is off; the emulated-warp test covers both.
#endif

constexpr unsigned kFullMask = 0xFFFFFFFFu;

__device__ __forceinline__ int warp_lane() { return (int)(threadIdx.x & 31); }

template <int DT>
__device__ __forceinline__ void broadcast_ray(const SceneDev &s, int src, const float *o, const float *dir, Skip skip,
                                              float *bo, float *bd, Skip &bskip) {
    const int D = NTR_D(DT, s);
    NTR_UNROLL
    for (int i = 0; i < D; ++i) { bo[i] = __shfl_sync(kFullMask, o[i], src); bd[i] = __shfl_sync(kFullMask, dir[i], src); }
    bskip.ref = __shfl_sync(kFullMask, skip.ref, src);
    bskip.lane = __shfl_sync(kFullMask, skip.lane, src);
}

// Per-warp count of lanes that have finished the traversal in progress (shared memory, one int per warp).
__device__ __forceinline__ void done_reset(int *done_ctr, bool enabled) {
    const unsigned idle = __ballot_sync(kFullMask, !enabled);
    if (warp_lane() == 0) *done_ctr = __popc(idle);
    __syncwarp();
}
__device__ __forceinline__ bool idle_lanes_wait(const int *done_ctr) {
    return *(const volatile int *)done_ctr >= NTR_COOP_MIN_DONE;
}

// true when serving the parked leaves cooperatively costs less than every parked lane scanning its own
__device__ __forceinline__ bool coop_pays(bool parked, uint32_t size) {
    const unsigned chunks = __reduce_add_sync(kFullMask, parked ? (size + 31u) / 32u * (unsigned)NTR_COOP_CHUNK_COST + (unsigned)NTR_COOP_OVERHEAD : 0u);
    const unsigned longest = __reduce_max_sync(kFullMask, parked ? size : 0u);
    return chunks < longest;
}

// ---- cooperative leaf, opaque variant ---------------------------------------------------------------------------
// kd_leaf::intersects for all-opaque simplex scenes (see leaf_opaque): nearest hit, first tested wins ties
// (tracer.hpp:1041-1082).  Lane j tests items j, j+32, ...; after every chunk the lanes share the smallest t so far as
// their cutoff; the winner is the smallest (t, item index).  Returns the owner's result on lane `src`.
template <int DT, int FLAGS>
__device__ __forceinline__ bool coop_leaf_opaque(const SceneDev &s, int src, uint32_t first, uint32_t size, const float *o,
                                                 const float *dir, Skip skip, HitRec &oh, Counters &cnt) {
    const int lane = warp_lane();
    float bo[DimCap<DT>::value], bd[DimCap<DT>::value];
    Skip bskip;
    broadcast_ray<DT>(s, src, o, dir, skip, bo, bd, bskip);
    const uint2 *items = s.leaf_items + first;
    float cut = __shfl_sync(kFullMask, oh.dist, src);       // shared cutoff: never above the owner's running nearest hit
    float my_t = FLT_MAX;                                   // this lane's own best
    uint32_t my_key = 0xFFFFFFFFu;                          // (item index << 7) | (batch lane + 1)
    for (uint32_t base = 0; base < size; base += 32) {
        const uint32_t k = base + (uint32_t)lane;
        if (k < size) {
            const uint2 it = lditem(items + k);
            uint32_t meta;
            if ((it.x >> 30) == NTR_REF_BATCH) {
                int index = bskip.ref == it.x ? bskip.lane : -1;
                const float dist = batch_test<DT, FLAGS>(s, it.y, bo, bd, index, cut, meta, cnt);
                if (dist) { my_t = dist; my_key = (k << 7) | (uint32_t)(index + 1); }
            } else if (it.x != bskip.ref) {
                if (FLAGS & NTR_F_COUNT) cnt.simplex_tests++;
                const float dist = simplex_single<DT>(s, it.y, bo, bd, cut, meta);
                if (dist) { my_t = dist; my_key = k << 7; }
            }
        }
        // hit distances are positive floats: their bit patterns order like the values
        cut = fminf(cut, u2f(__reduce_min_sync(kFullMask, f2u(my_t))));
    }
    const uint32_t best_t = __reduce_min_sync(kFullMask, my_key != 0xFFFFFFFFu ? f2u(my_t) : 0xFFFFFFFFu);
    const uint32_t best_key = __reduce_min_sync(kFullMask, (my_key != 0xFFFFFFFFu && f2u(my_t) == best_t) ? my_key : 0xFFFFFFFFu);
    if (best_key == 0xFFFFFFFFu) return false;
    if (lane == src) {
        oh.dist = u2f(best_t);
        oh.ref = lditem(items + (best_key >> 7)).x;
        oh.lane = (int)(best_key & 127u) - 1;
    }
    return true;
}

// ---- cooperative leaf, general variant --------------------------------------------------------------------------
// kd_leaf<Store,true>::intersects (tracer.hpp:977-1086) for one ray, evaluated by the whole warp: lane j tests item
// base+j of every 32-item chunk against the owner's cutoff at the start of the chunk (prim_eval), the owner replays
// the results in leaf order (replay_item).  Lanes never consult the owner's mailbox: an item the owner would skip is
// evaluated for nothing and dropped by the replay.  While the owner's mailbox can still answer (at most
// NTR_MAILBOX_CAP entries: the first two chunks of a traversal at most) every item is replayed, misses included;
// once it has switched itself off a plain miss has exactly one effect -- it is "the last test" whose result
// the final trim uses (Q3) -- so only the hits and partial writes are replayed and the gaps are settled with a mask.
template <int DT, int FLAGS>
__device__ __forceinline__ bool coop_leaf_general(const SceneDev &s, int src, uint32_t first, uint32_t size, const float *o,
                                                  const float *dir, Skip skip, HitRec &oh, GenState<DT> &g, Counters &cnt) {
    const int D = NTR_D(DT, s);
    const int lane = warp_lane();
    float bo[DimCap<DT>::value], bd[DimCap<DT>::value];
    Skip bskip;
    broadcast_ray<DT>(s, src, o, dir, skip, bo, bd, bskip);
    const uint2 *items = s.leaf_items + first;
    const int h_start = g.th.n;             // meaningful on the owner only
    float dist = 0;
    bool phase1 = false;
    // Exact mailbox (scenes with big leaves): every lane looks its item up in the OWNER's column of the scene-wide table
    // and, after the tests, the lanes enter their items there -- lanes whose items share a word are found with a
    // match-any vote and the lowest of them writes the word once.  The owner's replay then skips the mailbox.
    const bool exact = g.mb.exact();                // the same on every lane (a property of the scene)
    MailboxStore owner_mb = {};
    if (exact) {
        const unsigned long long colp = (unsigned long long)(size_t)g.mb.big->col;
        const uint32_t lo32 = __shfl_sync(kFullMask, (uint32_t)colp, src), hi32 = __shfl_sync(kFullMask, (uint32_t)(colp >> 32), src);
        owner_mb = *g.mb.big;
        owner_mb.col = (uint32_t *)(size_t)(((unsigned long long)hi32 << 32) | lo32);
        owner_mb.gen = __shfl_sync(kFullMask, g.mb.big->gen, src);
    }
    for (uint32_t base = 0; base < size; base += 32) {
        const uint32_t n = size - base < 32u ? size - base : 32u;
        const float cutoff0 = __shfl_sync(kFullMask, oh.dist, src);
        const bool mailbox_on = !exact && __shfl_sync(kFullMask, (int)(g.mb.n <= NTR_MAILBOX_CAP), src) != 0;
        ChunkEval<DT> e;
        e.dist = 0; e.wmask = 0; e.meta = 0; e.lane = -1; e.skipped = false; e.geom = false;
    NTR_UNROLL
        for (int i = 0; i < D; ++i) { e.P[i] = 0; e.N[i] = 0; }
        bool tested = false;
        uint32_t my_item = NTR_NONE_REF;
        if ((uint32_t)lane < n) {
            const uint2 it = lditem(items + base + lane);
            // the primitive the ray leaves from is skipped by identity, never evaluated
            if (!(((it.x >> 30) != NTR_REF_BATCH) && it.x == bskip.ref) && !(exact && MailboxStore::has(s, owner_mb.col, owner_mb.gen, it.x))) {
                prim_eval<DT, FLAGS>(s, it, bo, bd, cutoff0, bskip, e, cnt);
                tested = true;
                my_item = it.x;
            }
        }
        unsigned m = __ballot_sync(kFullMask, e.dist != 0 || e.wmask != 0);        // hits and partial writes
        const unsigned tm = __ballot_sync(kFullMask, tested);
        if (exact && tm) {
            const uint32_t key = tested ? MailboxStore::key_of(s, my_item) : 0u;
            const uint32_t word = tested ? key / NTR_MAILBOX_BITS_PER_WORD : 0x80000000u | (uint32_t)lane;    // idle lanes: a group of one
            const uint32_t bit = tested ? 1u << (key % NTR_MAILBOX_BITS_PER_WORD) : 0u;
            const unsigned peers = __match_any_sync(kFullMask, word);
            const int rounds = (int)__reduce_max_sync(kFullMask, (unsigned)__popc(peers));
            unsigned rem = peers;
            uint32_t bits = 0;
            for (int r = 0; r < rounds; ++r) {                  // OR of the bits of the lanes that share this lane's word
                const int j = rem ? __ffs(rem) - 1 : lane;
                rem &= rem - 1;
                bits |= __shfl_sync(kFullMask, bit, j);
            }
            if (tested && lane == __ffs(peers) - 1) {
                uint32_t *p = owner_mb.col + (size_t)word * s.mb_threads;
                const uint32_t w = *p;
                *p = ((w >> 24) == owner_mb.gen ? w : owner_mb.gen << 24) | bits;
            }
            __syncwarp();
        }
        uint32_t prev = 0;
        for (;;) {
            const uint32_t j = m ? (uint32_t)(__ffs(m) - 1) : n;
            if (lane == src && j > prev) {
                if (mailbox_on) {           // plain misses up to the next interesting item: the owner replays them alone
                    for (uint32_t k = prev; k < j; ++k) {
                        ChunkEval<DT> miss;
                        miss.dist = 0; miss.wmask = 0; miss.meta = 0; miss.lane = -1; miss.skipped = false; miss.geom = true;
                        replay_item<DT, FLAGS>(s, lditem(items + base + k), o, dir, skip, oh, g, cnt, phase1, dist, miss);
                    }
                } else {
                    const unsigned gap = (j >= 32 ? 0xFFFFFFFFu : ((1u << j) - 1u)) & ~((1u << prev) - 1u);
                    if (tm & gap) dist = 0;                     // some item of the gap was tested and missed
                }
            }
            if (j >= n) break;
            ChunkEval<DT> r;
            r.dist = __shfl_sync(kFullMask, e.dist, j); r.lane = __shfl_sync(kFullMask, e.lane, j);
            r.wmask = __shfl_sync(kFullMask, e.wmask, j); r.meta = __shfl_sync(kFullMask, e.meta, j);
            r.geom = __shfl_sync(kFullMask, (int)e.geom, j) != 0;
            r.skipped = false;
            if (FLAGS & NTR_F_GENERAL) {
                // only solids carry geometry out of the evaluation (warp-uniform branch: r.geom is the same on all lanes)
                if (r.geom) {
    NTR_UNROLL
                    for (int i = 0; i < D; ++i) { r.P[i] = __shfl_sync(kFullMask, e.P[i], j); r.N[i] = __shfl_sync(kFullMask, e.N[i], j); }
                }
            }
            if (lane == src) replay_item<DT, FLAGS>(s, lditem(items + base + j), o, dir, skip, oh, g, cnt, phase1, dist, r, exact);
            prev = j + 1;
            m &= m - 1;
        }
    }
    if (lane != src || !phase1) return false;
    g.th.trim(dist, h_start);
    return true;
}

// ---- nearest-hit traversal of a warp ----------------------------------------------------------------------------
// The state machine of trace_nearest (same frames, same unwinding rules) per lane, with parking at big leaves.
enum : int { NTR_S_DESCEND = 0, NTR_S_LEAF = 1, NTR_S_PARKED = 2, NTR_S_UNWIND = 3, NTR_S_DONE = 4 };

template <int DT, int FLAGS>
__device__ __forceinline__ bool trace_nearest_warp(const SceneDev &s, bool enabled, const float *o, const float *dir,
                                                   Skip skip, float t_near, float t_far, HitRec &oh, GenState<DT> &g,
                                                   Counters &cnt, int *done_ctr) {
    const int lane = warp_lane();
    done_reset(done_ctr, enabled);
    RaySlab<DT> rs;
    rs.init(s, dir);
    const float *invdir = rs.invdir;
    TravStack st;
    int sp = 0;
    uint32_t node = s.root;
    MiniMailbox mm;
    if (FLAGS & NTR_F_GENERAL) { g.mb.clear(); } else { mm.clear(); }
    int state = enabled ? NTR_S_DESCEND : NTR_S_DONE;
    bool result = false, ret = false;
    uint4 leaf = make_uint4(0u, 0u, 0u, 0u);
    for (;;) {
        // ---- per lane: until the ray is finished or parked at a big leaf ----
        while (state != NTR_S_DONE && state != NTR_S_PARKED) {
            if (state == NTR_S_DESCEND) {
                result = false;
                state = NTR_S_UNWIND;                       // falling off the tree is a miss
                while (node != NTR_NULL_NODE) {
                    const uint4 n = ldnode(s.nodes + node);
                    if (n.x & NTR_LEAF_FLAG) {
                        leaf = n;
                        state = (n.z >= (uint32_t)NTR_COOP_LEAF_MIN && idle_lanes_wait(done_ctr)) ? NTR_S_PARKED : NTR_S_LEAF;
                        break;
                    }
                    if (FLAGS & NTR_F_COUNT) cnt.node_steps++;
                    const int axis = (int)n.x;
                    const float split = u2f(n.y);
                    const float da = vsel<DT>(dir, axis), oa = vsel<DT>(o, axis);
                    if (da != 0) {
                        if (oa == split) { node = da > 0 ? n.w : n.z; continue; }
                        const float t = (split - oa) * vsel<DT>(invdir, axis);
                        const uint32_t n_near = oa > split ? n.w : n.z;
                        const uint32_t n_far = oa > split ? n.z : n.w;
                        if (t < 0 || t > t_far) { node = n_near; continue; }
                        if (t < t_near) { node = n_far; continue; }
                        if (n_near != NTR_NULL_NODE) {
                            if (n_far == NTR_NULL_NODE) { node = n_near; t_far = t; continue; }   // `|| !n_far) return hit`
                            if (sp < NTR_STACK_CAP) {
                                st.node[sp] = n_far; st.t[sp] = t; st.t_far[sp] = t_far;
                                if (FLAGS & NTR_F_GENERAL) st.h_start[sp] = (unsigned char)g.th.n;
                                ++sp;
                            }
                            node = n_near;
                            t_far = t;
                            continue;
                        }
                        node = n_far;
                        t_near = t;
                        continue;
                    }
                    node = oa >= split ? n.w : n.z;
                }
            } else if (state == NTR_S_LEAF) {
                if (FLAGS & NTR_F_GENERAL) result = leaf_general<DT, FLAGS>(s, leaf, o, dir, rs, skip, oh, g, cnt);
                else result = leaf_opaque<DT, FLAGS>(s, leaf, o, dir, rs, skip, oh, mm, cnt);
                state = NTR_S_UNWIND;
            } else {                                        // NTR_S_UNWIND (see trace_nearest)
                for (;;) {
                    if (sp == 0) { ret = result; state = NTR_S_DONE; atomicAdd(done_ctr, 1); break; }
                    --sp;
                    const uint32_t fnode = st.node[sp];
                    if (fnode == NTR_FRAME_AFTER_FAR) {
                        if (FLAGS & NTR_F_GENERAL) { if (result) g.th.trim(oh.dist, st.h_start[sp]); }
                        result = true;
                        continue;
                    }
                    const float t = st.t[sp];
                    if (result && oh.dist <= t) continue;                   // tracer.hpp:1214
                    node = fnode;
                    t_near = t;
                    t_far = st.t_far[sp];
                    if (result) {                                           // tracer.hpp:1216-1231
                        st.node[sp] = NTR_FRAME_AFTER_FAR;                  // h_start[sp] stays
                        ++sp;
                    }
                    state = NTR_S_DESCEND;
                    break;
                }
            }
        }
        // ---- the warp: serve the parked rays ----
        unsigned pm = __ballot_sync(kFullMask, state == NTR_S_PARKED);
        if (!pm) break;                                     // nobody parked: every lane is done
        if (!coop_pays(state == NTR_S_PARKED, leaf.z)) {
            if (state == NTR_S_PARKED) state = NTR_S_LEAF;  // coherent rays at the same big leaf: scan per lane
            continue;
        }
        while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1;
            const uint32_t first = __shfl_sync(kFullMask, leaf.y, src), size = __shfl_sync(kFullMask, leaf.z, src);
            bool r;
            if (FLAGS & NTR_F_GENERAL) r = coop_leaf_general<DT, FLAGS>(s, src, first, size, o, dir, skip, oh, g, cnt);
            else r = coop_leaf_opaque<DT, FLAGS>(s, src, first, size, o, dir, skip, oh, cnt);
            if (lane == src) { result = r; state = NTR_S_UNWIND; }
        }
    }
    return ret;
}

// ---- occlusion traversal of a warp ------------------------------------------------------------------------------
// kd_leaf::occludes (tracer.hpp:1088-1124) for one ray by the whole warp: any opaque hit nearer than the light
// decides (the transparent blockers collected so far are never looked at then, light_reaches returns false);
// otherwise the transparent blockers of the chunk are appended in leaf order.
template <int DT, int FLAGS>
__device__ __forceinline__ bool coop_leaf_occludes(const SceneDev &s, int src, uint32_t first, uint32_t size, const float *o,
                                                   const float *dir, float ldistance, Skip skip, HitList *hits, Counters &cnt) {
    const int lane = warp_lane();
    float bo[DimCap<DT>::value], bd[DimCap<DT>::value];
    Skip bskip;
    broadcast_ray<DT>(s, src, o, dir, skip, bo, bd, bskip);
    const float bld = __shfl_sync(kFullMask, ldistance, src);
    const uint2 *items = s.leaf_items + first;
    for (uint32_t base = 0; base < size; base += 32) {
        const uint32_t k = base + (uint32_t)lane;
        float dist = 0;
        uint32_t meta = 0;
        int hl = -1;
        if (k < size) {
            const uint2 it = lditem(items + k);
            const uint32_t kind = it.x >> 30;
            if (kind == NTR_REF_BATCH) {
                hl = bskip.ref == it.x ? bskip.lane : -1;
                dist = batch_test<DT, FLAGS>(s, it.y, bo, bd, hl, bld, meta, cnt);
            } else if (it.x != bskip.ref) {
                if (kind == NTR_REF_SIMPLEX) {
                    if (FLAGS & NTR_F_COUNT) cnt.simplex_tests++;
                    dist = simplex_single<DT>(s, it.y, bo, bd, bld, meta);
                } else if (FLAGS & NTR_F_GENERAL) {
                    float P[DimCap<DT>::value], N[DimCap<DT>::value];
                    uint32_t wmask;
                    if (FLAGS & NTR_F_COUNT) cnt.solid_tests++;
                    dist = solid_test<DT>(s, it.x & NTR_IDX_MASK, bo, bd, bld, P, N, wmask, meta);
                }
            }
        }
        const bool hit = dist != 0;
        if (!(FLAGS & NTR_F_GENERAL)) {
            if (__any_sync(kFullMask, hit)) return true;
            continue;
        }
        if (__any_sync(kFullMask, hit && (meta & NTR_META_OPAQUE))) return true;
        unsigned tr = __ballot_sync(kFullMask, hit);
        while (tr) {
            const int j = __ffs(tr) - 1;
            tr &= tr - 1;
            const float d = __shfl_sync(kFullMask, dist, j);
            const int l = __shfl_sync(kFullMask, hl, j);
            if (lane == src) hits->add(d, lditem(items + base + j).x, l);
        }
    }
    return false;
}

// _occludes (tracer.hpp:1258-1307) per lane with parking at big leaves; see trace_occludes for the frames.
template <int DT, int FLAGS>
__device__ __forceinline__ bool trace_occludes_warp(const SceneDev &s, bool enabled, const float *o, const float *dir,
                                                    float ldistance, Skip skip, float t_near, float t_far, HitList *hits,
                                                    Counters &cnt, int *done_ctr) {
    const int lane = warp_lane();
    done_reset(done_ctr, enabled);
    RaySlab<DT> rs;
    rs.init(s, dir);
    const float *invdir = rs.invdir;
    uint32_t st_node[NTR_STACK_CAP];
    float st_t[NTR_STACK_CAP], st_tfar[NTR_STACK_CAP];
    int sp = 0;
    uint32_t node = s.root;
    int state = enabled ? NTR_S_DESCEND : NTR_S_DONE;
    bool ret = false;
    uint4 leaf = make_uint4(0u, 0u, 0u, 0u);
    for (;;) {
        while (state != NTR_S_DONE && state != NTR_S_PARKED) {
            if (state == NTR_S_DESCEND) {
                state = NTR_S_UNWIND;
                while (node != NTR_NULL_NODE) {
                    const uint4 n = ldnode(s.nodes + node);
                    if (n.x & NTR_LEAF_FLAG) {
                        leaf = n;
                        state = (n.z >= (uint32_t)NTR_COOP_LEAF_MIN && idle_lanes_wait(done_ctr)) ? NTR_S_PARKED : NTR_S_LEAF;
                        break;
                    }
                    if (FLAGS & NTR_F_COUNT) cnt.node_steps++;
                    const int axis = (int)n.x;
                    const float split = u2f(n.y);
                    const float da = vsel<DT>(dir, axis), oa = vsel<DT>(o, axis);
                    if (da != 0) {
                        if (oa == split) { node = da > 0 ? n.w : n.z; continue; }
                        const float t = (split - oa) * vsel<DT>(invdir, axis);
                        const uint32_t n_near = oa > split ? n.w : n.z;
                        const uint32_t n_far = oa > split ? n.z : n.w;
                        if (t < 0 || t > t_far) { node = n_near; continue; }
                        if (t < t_near) { node = n_far; continue; }
                        if (n_near != NTR_NULL_NODE) {
                            if (n_far == NTR_NULL_NODE) { t_far = t; node = n_near; continue; }
                            if (sp < NTR_STACK_CAP) { st_node[sp] = n_far; st_t[sp] = t; st_tfar[sp] = t_far; ++sp; }
                            node = n_near;
                            t_far = t;
                            continue;
                        }
                        if (t < ldistance) break;           // near child null: falls to the same test (:1297-1298)
                        t_near = t;
                        node = n_far;
                        continue;
                    }
                    node = oa >= split ? n.w : n.z;
                }
            } else if (state == NTR_S_LEAF) {
                if (leaf_occludes<DT, FLAGS>(s, leaf, o, dir, rs, ldistance, skip, hits, cnt)) { ret = true; state = NTR_S_DONE; atomicAdd(done_ctr, 1); }
                else state = NTR_S_UNWIND;
            } else {
                // the (sub)call returned false: resume the innermost pending far child
                for (;;) {
                    if (sp == 0) { ret = false; state = NTR_S_DONE; atomicAdd(done_ctr, 1); break; }
                    --sp;
                    if (st_t[sp] < ldistance) continue;     // `return false` of that frame
                    node = st_node[sp];
                    t_near = st_t[sp];
                    t_far = st_tfar[sp];
                    state = NTR_S_DESCEND;
                    break;
                }
            }
        }
        unsigned pm = __ballot_sync(kFullMask, state == NTR_S_PARKED);
        if (!pm) break;
        if (!coop_pays(state == NTR_S_PARKED, leaf.z)) {
            if (state == NTR_S_PARKED) state = NTR_S_LEAF;
            continue;
        }
        while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1;
            const uint32_t first = __shfl_sync(kFullMask, leaf.y, src), size = __shfl_sync(kFullMask, leaf.z, src);
            const bool r = coop_leaf_occludes<DT, FLAGS>(s, src, first, size, o, dir, ldistance, skip, hits, cnt);
            if (lane == src) {
                if (r) { ret = true; state = NTR_S_DONE; atomicAdd(done_ctr, 1); }
                else state = NTR_S_UNWIND;
            }
        }
    }
    return ret;
}

// composite_scene::light_reaches (tracer.hpp:1750-1766) for the lanes with need = true
template <int DT, int FLAGS>
__device__ __forceinline__ bool light_reaches_warp(const SceneDev &s, bool need, const float *o, const float *dir,
                                                   float ldistance, Skip skip, float *filtered, Counters &cnt, int *done_ctr) {
    HitList hits;
    if (FLAGS & NTR_F_GENERAL) hits.clear();
    if (need) cnt.shadow_rays++;
    const bool occluded = trace_occludes_warp<DT, FLAGS>(s, need, o, dir, ldistance, skip, 0.0f, FLT_MAX, &hits, cnt, done_ctr);
    if (!need || occluded) return false;
    if (FLAGS & NTR_F_GENERAL) {
        if (hits.dropped) cnt.truncated++;
        if (hits.n) {
            hits.sort_and_unique();
            for (int i = hits.n - 1; i >= 0; --i) {
                const Mat m = load_mat(s, target_meta<DT>(s, hits.ref[i], hits.lane[i]));
                const float f = 1 - m.opacity;
                filtered[0] *= f; filtered[1] *= f; filtered[2] *= f;
            }
        }
    }
    return true;
}

// composite_scene::ray_color (tracer.hpp:1856-1883) for the 32 rays of a warp; see ray_color for the linearisation.
template <int DT, int FLAGS, typename EMIT>
__device__ __forceinline__ void ray_color_warp(const SceneDev &s, bool enabled, const float *o, const float *dir, int depth,
                                               Skip source, const float *weight, float *acc, EMIT &emit, Counters &cnt,
                                               HitRec *primary_out, int *done_ctr, MailboxStore *ms) {
    const int D = NTR_D(DT, s);
    GenState<DT> g;
    g.mb.big = (ms && ms->col) ? ms : nullptr;
    HitRec oh;
    oh.dist = FLT_MAX; oh.ref = NTR_NONE_REF; oh.lane = -1;
    if (FLAGS & NTR_F_GENERAL) {
        g.th.clear();
    NTR_UNROLL
        for (int i = 0; i < D; ++i) { g.hitP[i] = 0; g.hitN[i] = 0; }
    }
    const float t0 = enabled ? aabb_distance<DT>(s, o, dir) : -1.0f;
    const bool hit = trace_nearest_warp<DT, FLAGS>(s, t0 >= 0, o, dir, source, t0, FLT_MAX, oh, g, cnt, done_ctr);
    if (enabled && primary_out) { *primary_out = oh; if (!hit) { primary_out->ref = NTR_NONE_REF; primary_out->dist = 0; } }

    float w[3] = {weight[0], weight[1], weight[2]};
    int n_layers = 0;
    if (FLAGS & NTR_F_GENERAL) {
        if (enabled && g.th.dropped) cnt.truncated++;
        if (enabled && g.th.n) g.th.sort_and_unique();
        n_layers = enabled ? g.th.n : 0;
    }
    const int n_total = enabled ? n_layers + (hit ? 1 : 0) : 0;
#if !NTR_WARP_SHADE
    // per-lane shading (shadow rays traced by the lane that needs them), as in ray_color
    for (int i = 0; i < n_total; ++i) {
        uint32_t ref;
        int hl;
        float wl[3];
        float P[DimCap<DT>::value], N[DimCap<DT>::value];
        if (i < n_layers) {
            ref = g.th.ref[i];
            hl = g.th.lane[i];
            const float op = load_mat(s, target_meta<DT>(s, ref, hl)).opacity;
            hit_geometry<DT, FLAGS>(s, ref, hl, g.th.dist[i], o, dir, P, N);
            wl[0] = w[0] * op; wl[1] = w[1] * op; wl[2] = w[2] * op;
            w[0] *= 1 - op; w[1] *= 1 - op; w[2] *= 1 - op;
        } else {
            ref = oh.ref;
            hl = oh.lane;
            wl[0] = w[0]; wl[1] = w[1]; wl[2] = w[2];
            if (FLAGS & NTR_F_GENERAL) {
    NTR_UNROLL
                for (int k = 0; k < D; ++k) { P[k] = g.hitP[k]; N[k] = g.hitN[k]; }
            } else {
                hit_geometry<DT, FLAGS>(s, ref, hl, oh.dist, o, dir, P, N);
            }
        }
        Bounce<DT> b;
        if (shade_hit<DT, FLAGS>(s, dir, P, N, ref, hl, depth, wl, acc, b, cnt)) emit(b);
    }
#else
    const int n_lights = s.n_point + s.n_global;
    // layers near -> far (the surviving transparent hits, then the opaque hit), lights in order: warp-uniform loops
    for (int i = 0; __any_sync(kFullMask, i < n_total); ++i) {
        const bool have = i < n_total;
        uint32_t ref = NTR_NONE_REF;
        int hl = -1;
        float wl[3] = {0, 0, 0};
        float P[DimCap<DT>::value], N[DimCap<DT>::value];
    NTR_UNROLL
        for (int k = 0; k < D; ++k) { P[k] = 0; N[k] = 0; }
        Mat m = {};
        if (have) {
            if (i < n_layers) {
                ref = g.th.ref[i];
                hl = g.th.lane[i];
                m = load_mat(s, target_meta<DT>(s, ref, hl));
                hit_geometry<DT, FLAGS>(s, ref, hl, g.th.dist[i], o, dir, P, N);
                wl[0] = w[0] * m.opacity; wl[1] = w[1] * m.opacity; wl[2] = w[2] * m.opacity;
                w[0] *= 1 - m.opacity; w[1] *= 1 - m.opacity; w[2] *= 1 - m.opacity;
            } else {
                ref = oh.ref;
                hl = oh.lane;
                m = load_mat(s, target_meta<DT>(s, ref, hl));
                wl[0] = w[0]; wl[1] = w[1]; wl[2] = w[2];
                if (FLAGS & NTR_F_GENERAL) {
    NTR_UNROLL
                    for (int k = 0; k < D; ++k) { P[k] = g.hitP[k]; N[k] = g.hitN[k]; }      // as the reference left it (Q12)
                } else {
                    hit_geometry<DT, FLAGS>(s, ref, hl, oh.dist, o, dir, P, N);
                }
            }
            cnt.shaded_hits++;
        }
        const Skip src = {ref, hl};
        ShadeAcc a = {{0, 0, 0}, {0, 0, 0}, 0};
        for (int li = 0; li < n_lights; ++li) {
            LightSample<DT> ls = {};
            int kind = NTR_LIGHT_NONE;
            if (have) kind = light_prepare<DT>(s, li, P, N, ls);
            float filtered[3] = {0, 0, 0};
            if (kind != NTR_LIGHT_NONE) { filtered[0] = ls.lc[0]; filtered[1] = ls.lc[1]; filtered[2] = ls.lc[2]; }
            if (s.shadows) {            // scene constant: uniform
                const bool need = kind == NTR_LIGHT_SHADOWED;
                if (__any_sync(kFullMask, need)) {
                    const bool reaches = light_reaches_warp<DT, FLAGS>(s, need, P, ls.lv, ls.dist, src, filtered, cnt, done_ctr);
                    if (need && !reaches) kind = NTR_LIGHT_NONE;
                }
            }
            if (kind != NTR_LIGHT_NONE) light_apply<DT>(s, kind, ls, filtered, m, dir, N, a);
        }
        if (have) {
            Bounce<DT> b;
            if (shade_finish<DT>(s, m, dir, P, N, src, depth, wl, acc, a, b, cnt)) emit(b);
        }
    }
#endif
    if (enabled && !hit) {
        const float I = vsel<DT>(dir, s.bg_axis);       // tracer.hpp:1866-1867
    NTR_UNROLL
        for (int c = 0; c < 3; ++c) {
            const float bg = I >= 0 ? s.bg1[c] * I + s.bg2[c] * (1 - I) : s.bg3[c] * -I + s.bg2[c] * (1 + I);
            acc[c] += w[c] * bg;
        }
    }
}

}  // namespace ntr
#endif  // NTR_WARP_CODE
