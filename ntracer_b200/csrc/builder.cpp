// builder.cpp -- host-side scene construction next to the render path:
//   * ntr_simplex_from_points: TrianglePrototype / Triangle.from_points for many simplexes at once
//     (reference src/tracer.hpp:442-462: generalized cross products, src/geometry.hpp:858-893)
//   * ntr_build_kdtree: k-d tree over axis-aligned item bounds, replacing build_kdtree
//     (reference src/tracer.hpp:1930-2455) for the GPU backend.
// Written from scratch: binned surface-area heuristic over all axes, items straddling the split plane go to both
// children, empty children become null (like kd_branch's null pointers), subtrees built by a small thread pool.
// It is NOT a port of the reference builder (no primitive clipping by separating-axis tests, no greedy SIMD batch
// grouping -- the reference's O(N^2) group_primitives is what makes it unusable at config-5 scale).  Any correct tree
// gives the same nearest hits; only shadows-on images depend on the topology.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <memory>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/ntracer_b200.h"
#include "errors.h"

namespace {

// ---- generalized cross product ---------------------------------------------------------------------------
// determinant of an m x m matrix (row-major, destroyed) by LU with partial pivoting
double det_inplace(double *a, int m) {
    double det = 1;
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int r = c + 1; r < m; ++r) if (fabs(a[r * m + c]) > fabs(a[piv * m + c])) piv = r;
        if (a[piv * m + c] == 0) return 0;
        if (piv != c) { for (int k = 0; k < m; ++k) std::swap(a[piv * m + k], a[c * m + k]); det = -det; }
        det *= a[c * m + c];
        for (int r = c + 1; r < m; ++r) {
            const double f = a[r * m + c] / a[c * m + c];
            if (f != 0) for (int k = c; k < m; ++k) a[r * m + k] -= f * a[c * m + k];
        }
    }
    return det;
}

// r = cross(vs[0..D-2]) with the reference's sign convention (geometry.hpp:858-871):
// r[i] = f_i * det(minor without component i), f_0 = +1 for odd D, -1 for even D, alternating
void cross(int D, const double *vs /*(D-1) x D*/, double *r) {
    const int m = D - 1;
    std::vector<double> tmp((size_t)m * m);
    double f = D % 2 ? 1.0 : -1.0;
    for (int i = 0; i < D; ++i) {
        for (int j = 0; j < m; ++j) {
            for (int k = 0; k < i; ++k) tmp[(size_t)k * m + j] = vs[(size_t)j * D + k];
            for (int k = i + 1; k < D; ++k) tmp[(size_t)(k - 1) * m + j] = vs[(size_t)j * D + k];
        }
        r[i] = f * det_inplace(tmp.data(), m);
        f = -f;
    }
}

// ---- k-d tree ------------------------------------------------------------------------------------------------
struct Builder {
    int D;
    const float *lo, *hi;
    int max_depth, split_threshold;
    double c_trav, c_isect;
    static constexpr int kBins = 32;
    static constexpr int kForceSplit = 128;     // culled builds: cells with at least this many items are split regardless
    // optional (ntr_build_kdtree_culled): the simplexes behind every item -- item i owns simplexes
    // item_first[i] .. item_first[i+1]-1, each with its own bounds and its record (face_normal[D], d, p1[D],
    // edge_normals[D-1][D]) -- so that an item is handed to a child cell only if one of its simplexes can touch the cell.
    // A simplex is the part of its hyperplane where the D barycentric functions the ray test evaluates (area_i =
    // edge_normal_i . (p1 - x) >= 0 for i = 1..D-1 and 1 - sum area_i >= 0, tracer.hpp:411-440) are non-negative; all of
    // them and the plane function are affine, so their ranges over a box come from the box's extreme corners.  The cell
    // cannot touch the simplex if its bounds miss the cell, or the plane function has one sign on the whole cell, or one
    // barycentric function is negative on the whole cell: necessary conditions only (separating axes: the box axes, the
    // normal, the D facet directions), each with a relative margin, so an item whose geometry meets the cell is never
    // dropped.  Items without simplexes (solids) are kept by their bounds alone.
    const uint32_t *item_first = nullptr;
    const float *s_lo = nullptr, *s_hi = nullptr, *s_rec = nullptr;

    // range of  c + sum_k g[k] * x[k]  over the box, and a magnitude for the rounding margin
    static void affine_range(int D, const float *g, double c, const double *blo, const double *bhi, double &mn, double &mx, double &mag) {
        mn = mx = c;
        mag = fabs(c);
        for (int k = 0; k < D; ++k) {
            const double x0 = (double)g[k] * blo[k], x1 = (double)g[k] * bhi[k];
            mn += std::min(x0, x1);
            mx += std::max(x0, x1);
            mag += std::max(fabs(x0), fabs(x1));
        }
    }

    bool touches(uint32_t item, const double *blo, const double *bhi) const {
        if (!item_first) return true;
        const uint32_t a = item_first[item], b = item_first[item + 1];
        if (a == b) return true;
        const size_t S = (size_t)(D + 1) * D + 1;
        double gsum[NTR_MAX_DIM];
        for (uint32_t s = a; s < b; ++s) {
            const float *l = s_lo + (size_t)s * D, *h = s_hi + (size_t)s * D, *rec = s_rec + (size_t)s * S;
            bool in = true;
            for (int k = 0; k < D && in; ++k) {
                const double tol = 1e-6 * (fabs(blo[k]) + fabs(bhi[k]) + 1e-30);
                if ((double)l[k] > bhi[k] + tol || (double)h[k] < blo[k] - tol) in = false;
            }
            if (!in) continue;
            double mn, mx, mag;
            affine_range(D, rec, (double)rec[D], blo, bhi, mn, mx, mag);            // the hyperplane: normal . x + d
            if (mn > 1e-5 * mag || mx < -1e-5 * mag) continue;
            // area_i(x) = e_i . p1 - e_i . x; all of them >= 0 and their sum <= 1 on the simplex
            const float *p1 = rec + D + 1;
            double csum = 0;
            for (int k = 0; k < D; ++k) gsum[k] = 0;
            for (int i = 0; i < D - 1 && in; ++i) {
                const float *e = rec + 2 * D + 1 + (size_t)i * D;
                double c = 0;
                for (int k = 0; k < D; ++k) { c += (double)e[k] * (double)p1[k]; gsum[k] += e[k]; }
                csum += c;
                // range of e . x over the box; area = c - e . x
                double emn, emx, emag;
                affine_range(D, e, 0.0, blo, bhi, emn, emx, emag);
                const double eps = 1e-5 * (emag + fabs(c)) + 1e-5;
                if (c - emn < -eps) in = false;           // the largest area on the cell is negative
                if (c - emx > 1 + eps) in = false;        // the smallest area on the cell is above 1
            }
            if (!in) continue;
            {   // sum of the areas <= 1
                double emn = 0, emag = 0;
                for (int k = 0; k < D; ++k) {
                    const double x0 = gsum[k] * blo[k], x1 = gsum[k] * bhi[k];
                    emn += std::max(x0, x1);              // largest gsum . x  ->  smallest sum of areas
                    emag += std::max(fabs(x0), fabs(x1));
                }
                if (csum - emn > 1 + 1e-5 * (emag + fabs(csum)) + 1e-5) continue;
            }
            return true;
        }
        return false;
    }

    struct Node { uint32_t meta, w1, w2, w3; std::vector<uint32_t> items; };
    // subtrees are built into private vectors and spliced together afterwards
    struct Tree { std::vector<ntr_node> nodes; std::vector<uint32_t> refs; };

    double area(const double *e) const {        // (D-1)-measure of the box surface, up to a constant factor
        double prod = 1;
        for (int i = 0; i < D; ++i) prod *= std::max(e[i], 1e-12);
        double s = 0;
        for (int i = 0; i < D; ++i) s += prod / std::max(e[i], 1e-12);
        return s;
    }

    // returns the node index inside `t` (NTR_NULL_NODE for an empty set)
    uint32_t build(Tree &t, std::vector<uint32_t> &idx, const double *nlo, const double *nhi, int depth) const {
        const size_t n = idx.size();
        if (n == 0) return NTR_NULL_NODE;
        auto make_leaf = [&]() {
            const uint32_t me = (uint32_t)t.nodes.size();
            t.nodes.push_back(ntr_node{NTR_LEAF_FLAG, (uint32_t)t.refs.size(), (uint32_t)n, 0});
            for (uint32_t i : idx) t.refs.push_back(i);
            return me;
        };
        if ((int)n <= split_threshold || depth >= max_depth) return make_leaf();
        double ext[NTR_MAX_DIM];
        for (int i = 0; i < D; ++i) ext[i] = nhi[i] - nlo[i];
        const double parent_area = area(ext);
        double best_cost = c_isect * (double)n;
        int best_axis = -1;
        double best_split = 0;
        std::vector<uint32_t> hl(kBins + 1), hh(kBins + 1);
        double ax_cost[NTR_MAX_DIM], ax_split[NTR_MAX_DIM];
        for (int ax = 0; ax < D; ++ax) { ax_cost[ax] = 1e300; ax_split[ax] = 0; }
        for (int ax = 0; ax < D; ++ax) {
            if (!(ext[ax] > 0)) continue;
            std::fill(hl.begin(), hl.end(), 0u);
            std::fill(hh.begin(), hh.end(), 0u);
            const double scale = kBins / ext[ax];
            for (uint32_t i : idx) {
                // bin b covers [nlo + b*w, nlo + (b+1)*w); plane p (1..kBins-1) sits at nlo + p*w
                int bl = (int)floor((lo[(size_t)i * D + ax] - nlo[ax]) * scale);
                int bh = (int)ceil((hi[(size_t)i * D + ax] - nlo[ax]) * scale);
                bl = std::min(std::max(bl, 0), kBins);
                bh = std::min(std::max(bh, 0), kBins);
                ++hl[bl];       // the item starts in bin bl: it is left of every plane p > bl
                ++hh[bh];       // the item ends at bin edge bh: it is right of every plane p < bh
            }
            uint32_t n_left = 0;        // items whose lower bound lies before plane p  -> left child
            uint32_t cum_h = 0;         // items whose upper bound lies at or before plane p-1
            double e2[NTR_MAX_DIM];
            memcpy(e2, ext, sizeof(double) * D);
            for (int p = 1; p < kBins; ++p) {
                n_left += hl[p - 1];                // lo-bin < p
                cum_h += hh[p - 1];
                const uint32_t ended = cum_h + hh[p];          // upper bound <= plane p: not in the right child
                const uint32_t n_right = (uint32_t)n - ended;
                const double w = ext[ax] * p / kBins;
                e2[ax] = w;
                const double al = area(e2);
                e2[ax] = ext[ax] - w;
                const double ar = area(e2);
                const double cost = c_trav + c_isect * (al * n_left + ar * n_right) / parent_area;
                if (cost < best_cost && (n_left < n || n_right < n)) {
                    best_cost = cost; best_axis = ax; best_split = nlo[ax] + w;
                }
                if (cost < ax_cost[ax]) { ax_cost[ax] = cost; ax_split[ax] = nlo[ax] + w; }
            }
        }
        if (item_first) {
            // With the simplexes at hand the counts of the sweep (bounding boxes) overestimate what the children will
            // hold, and the sweep gives up splitting too early.  Re-evaluate a few candidate planes -- per axis the
            // sweep's best and the spatial median -- with the children's real contents and take the cheapest.
            best_cost = c_isect * (double)n;
            best_axis = -1;
            // A cell that still holds hundreds of items is split even when no plane pays by the greedy estimate (the centre
            // of a star polytope: every plane through it leaves most of the big simplexes on both sides, and only several
            // levels further down do the cells get small enough to lose them): at the cheapest plane that sends fewer than
            // all items to each side.  Leaves of that size are what a frame's slowest rays walk.
            double forced_cost = 1e300, forced_split = 0;
            int forced_axis = -1;
            double l_hi[NTR_MAX_DIM], r_lo[NTR_MAX_DIM], e2[NTR_MAX_DIM];
            for (int ax = 0; ax < D; ++ax) {
                if (!(ext[ax] > 0)) continue;
                for (int cand = 0; cand < 2; ++cand) {
                    const double sp = cand == 0 ? ax_split[ax] : nlo[ax] + 0.5 * ext[ax];
                    if (cand == 0 && !(ax_cost[ax] < 1e300)) continue;
                    const float split = (float)sp;
                    if (!((double)split > nlo[ax] && (double)split < nhi[ax])) continue;
                    memcpy(l_hi, nhi, sizeof(double) * D);
                    memcpy(r_lo, nlo, sizeof(double) * D);
                    l_hi[ax] = split; r_lo[ax] = split;
                    size_t nl = 0, nr = 0;
                    for (uint32_t i : idx) {
                        const float a = lo[(size_t)i * D + ax], b = hi[(size_t)i * D + ax];
                        const bool flat = a == split && b == split;
                        if ((a < split || flat) && touches(i, nlo, l_hi)) ++nl;
                        if ((b > split || flat) && touches(i, r_lo, nhi)) ++nr;
                    }
                    if (nl == n && nr == n) continue;
                    memcpy(e2, ext, sizeof(double) * D);
                    e2[ax] = (double)split - nlo[ax];
                    const double al = area(e2);
                    e2[ax] = nhi[ax] - (double)split;
                    const double ar = area(e2);
                    const double cost = c_trav + c_isect * (al * (double)nl + ar * (double)nr) / parent_area;
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = split; }
                    if (cost < forced_cost && nl < n && nr < n) { forced_cost = cost; forced_axis = ax; forced_split = split; }
                }
            }
            if (best_axis < 0 && n >= (size_t)kForceSplit && forced_axis >= 0) { best_axis = forced_axis; best_split = forced_split; }
        }
        if (best_axis < 0) return make_leaf();
        const float split = (float)best_split;
        std::vector<uint32_t> li, ri;
        li.reserve(n); ri.reserve(n);
        double l_hi[NTR_MAX_DIM], r_lo[NTR_MAX_DIM];
        memcpy(l_hi, nhi, sizeof(double) * D);
        memcpy(r_lo, nlo, sizeof(double) * D);
        l_hi[best_axis] = split;
        r_lo[best_axis] = split;
        for (uint32_t i : idx) {
            const float a = lo[(size_t)i * D + best_axis], b = hi[(size_t)i * D + best_axis];
            const bool flat = a == split && b == split;
            if ((a < split || flat) && touches(i, nlo, l_hi)) li.push_back(i);
            if ((b > split || flat) && touches(i, r_lo, nhi)) ri.push_back(i);
        }
        if (li.size() == n && ri.size() == n) return make_leaf();
        std::vector<uint32_t>().swap(idx);                  // free before recursing
        const uint32_t me = (uint32_t)t.nodes.size();
        uint32_t bits;
        memcpy(&bits, &split, 4);
        t.nodes.push_back(ntr_node{(uint32_t)best_axis, bits, NTR_NULL_NODE, NTR_NULL_NODE});
        const uint32_t l = build(t, li, nlo, l_hi, depth + 1);
        const uint32_t r = build(t, ri, r_lo, nhi, depth + 1);
        if (l == NTR_NULL_NODE && r == NTR_NULL_NODE) { t.nodes.pop_back(); return NTR_NULL_NODE; }
        t.nodes[me].w2 = l;
        t.nodes[me].w3 = r;
        return me;
    }
};

}  // namespace

extern "C" {

NTR_API void ntr_free(void *p) { free(p); }

// points: n x D x D (n simplexes of D vertices); records: n x ((D+1)*D+1) = face_normal[D], d, p1[D], edge_normals[D-1][D]
NTR_API int ntr_simplex_from_points(int dim, uint32_t n, const float *points, float *records) {
    if (dim < 3 || dim > NTR_MAX_DIM) return ntr_fail(NTR_ERR_VALUE, "dimension must be between 3 and %d", NTR_MAX_DIM);
    if ((!points && n) || (!records && n)) return ntr_fail(NTR_ERR_VALUE, "points / records is NULL");
    const int D = dim, S = (D + 1) * D + 1;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nthreads = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, n / 256));
    auto work = [&](uint32_t a, uint32_t b) {
        std::vector<double> vs((size_t)(D - 1) * D), tmp((size_t)(D - 1) * D), N(D), r(D);
        for (uint32_t s = a; s < b; ++s) {
            const float *P = points + (size_t)s * D * D;
            float *rec = records + (size_t)s * S;
            for (int j = 0; j < D - 1; ++j)
                for (int k = 0; k < D; ++k) vs[(size_t)j * D + k] = (double)P[(size_t)(j + 1) * D + k] - (double)P[k];
            cross(D, vs.data(), N.data());
            double sq = 0;
            for (int k = 0; k < D; ++k) sq += N[k] * N[k];
            double dot = 0;
            for (int k = 0; k < D; ++k) { rec[k] = (float)N[k]; dot += (double)rec[k] * (double)P[k]; rec[D + 1 + k] = P[k]; }
            rec[D] = (float)-dot;                                   // recalculate_d, tracer.hpp:472-474
            for (int i = 0; i < D - 1; ++i) {                       // tracer.hpp:454-461
                tmp = vs;
                for (int k = 0; k < D; ++k) tmp[(size_t)i * D + k] = N[k];
                cross(D, tmp.data(), r.data());
                for (int k = 0; k < D; ++k) rec[2 * D + 1 + (size_t)i * D + k] = (float)(r[k] / sq);
            }
        }
    };
    std::vector<std::thread> th;
    uint32_t started = 0;                   // simplexes handed to threads so far
    try {
        for (unsigned t = 0; t + 1 < nthreads; ++t) {
            const uint32_t a = (uint32_t)((uint64_t)n * t / nthreads), b = (uint32_t)((uint64_t)n * (t + 1) / nthreads);
            th.emplace_back(work, a, b);
            started = b;
        }
    } catch (const std::system_error &) {}  // no more threads to be had: the caller's thread does the rest
    int rc = NTR_OK;
    try { work(started, n); } catch (const std::bad_alloc &) { rc = ntr_fail(NTR_ERR_MEMORY, "out of memory"); }
    for (auto &t : th) t.join();
    return rc;
}

// Groups n items (simplexes, by their bounds lo/hi: n x D) into runs of `group` spatially close items: order_out receives a
// permutation of 0..n-1 in which every consecutive run of `group` entries is one group (what becomes a triangle_batch);
// the last n % group entries are the left-overs that stay single primitives.  The reference groups greedily in O(n^2)
// (group_primitives, src/tracer.hpp:2395-2427: each unassigned triangle takes the v_real::size - 1 nearest unassigned
// ones by grouping_metric); this is an O(n log n) recursive median split of the item centres along their widest axis,
// with every split placed on a multiple of `group`.
NTR_API int ntr_group_items(int dim, uint32_t n, const float *lo, const float *hi, int group, uint32_t *order_out) {
    if (dim < 3 || dim > NTR_MAX_DIM) return ntr_fail(NTR_ERR_VALUE, "dimension must be between 3 and %d", NTR_MAX_DIM);
    if (group < 1 || group > 64) return ntr_fail(NTR_ERR_VALUE, "group size must be in 1..64");
    if (n && (!lo || !hi || !order_out)) return ntr_fail(NTR_ERR_VALUE, "NULL argument");
    const int D = dim;
    try {
        std::vector<float> c((size_t)n * D);
        for (size_t k = 0; k < (size_t)n * D; ++k) c[k] = 0.5f * (lo[k] + hi[k]);
        for (uint32_t i = 0; i < n; ++i) order_out[i] = i;
        std::vector<std::pair<uint32_t, uint32_t>> stack;
        stack.push_back({0u, n});
        while (!stack.empty()) {
            const auto [a, b] = stack.back();
            stack.pop_back();
            if (b - a <= (uint32_t)group) continue;
            int axis = 0;
            float best = -1.0f;
            for (int k = 0; k < D; ++k) {
                float mn = c[(size_t)order_out[a] * D + k], mx = mn;
                for (uint32_t i = a + 1; i < b; ++i) {
                    const float v = c[(size_t)order_out[i] * D + k];
                    mn = std::min(mn, v); mx = std::max(mx, v);
                }
                if (mx - mn > best) { best = mx - mn; axis = k; }
            }
            const uint32_t groups = (b - a) / (uint32_t)group;              // whole groups in this range (>= 1)
            const uint32_t mid = a + std::max(1u, (groups + 1) / 2) * (uint32_t)group;
            if (mid >= b) continue;
            std::nth_element(order_out + a, order_out + mid, order_out + b, [&](uint32_t x, uint32_t y) {
                const float cx = c[(size_t)x * D + axis], cy = c[(size_t)y * D + axis];
                return cx < cy || (cx == cy && x < y);
            });
            stack.push_back({a, mid});
            stack.push_back({mid, b});
        }
    } catch (const std::bad_alloc &) {
        return ntr_fail(NTR_ERR_MEMORY, "out of memory grouping the items");
    }
    return NTR_OK;
}

// lo/hi: n x D item bounds.  Outputs are malloc'ed (release with ntr_free): nodes (16-byte ntr_node, root = node 0 or
// NTR_NULL_NODE when n == 0), refs = item indices per leaf (the caller turns them into leaf refs), boundary = 2 x D.
NTR_API int ntr_build_kdtree(int dim, uint32_t n, const float *lo, const float *hi, int max_depth, int split_threshold,
                             float traversal_cost, float intersection_cost, ntr_node **nodes_out, uint32_t *n_nodes_out,
                             uint32_t **refs_out, uint32_t *n_refs_out, uint32_t *root_out, float *boundary_out) {
    return ntr_build_kdtree_culled(dim, n, lo, hi, nullptr, 0, nullptr, nullptr, nullptr, max_depth, split_threshold, traversal_cost,
                                   intersection_cost, nodes_out, n_nodes_out, refs_out, n_refs_out, root_out, boundary_out);
}

// The same with the simplexes behind the items (see Builder::touches): what the reference's builder does with its exact
// overlap tests (src/tracer.hpp:1465-1675, 2284-2354), here as separating axes: bounds, hyperplane, facet directions.
NTR_API int ntr_build_kdtree_culled(int dim, uint32_t n, const float *lo, const float *hi, const uint32_t *item_first,
                                    uint32_t n_simplex, const float *s_lo, const float *s_hi, const float *s_records,
                                    int max_depth, int split_threshold, float traversal_cost, float intersection_cost,
                                    ntr_node **nodes_out, uint32_t *n_nodes_out, uint32_t **refs_out, uint32_t *n_refs_out,
                                    uint32_t *root_out, float *boundary_out) {
    if (item_first) {
        if (n_simplex && (!s_lo || !s_hi || !s_records)) return ntr_fail(NTR_ERR_VALUE, "simplex bounds / records are NULL");
        if (item_first[0] != 0) return ntr_fail(NTR_ERR_VALUE, "item_first[0] must be 0");
        for (uint32_t k = 0; k < n; ++k)
            if (item_first[k + 1] < item_first[k] || item_first[k + 1] > n_simplex)
                return ntr_fail(NTR_ERR_VALUE, "item_first is not a non-decreasing sequence within the simplex count");
        if (dim >= 3 && dim <= NTR_MAX_DIM)
            for (size_t k = 0; k < (size_t)n_simplex * ((size_t)(dim + 1) * dim + 1); ++k) {
                if (!(s_records[k] > -3e38f && s_records[k] < 3e38f))
                    return ntr_fail(NTR_ERR_VALUE, "simplex %zu: record holds NaN or infinity", k / ((size_t)(dim + 1) * dim + 1));
            }
    }
    if (dim < 3 || dim > NTR_MAX_DIM) return ntr_fail(NTR_ERR_VALUE, "dimension must be between 3 and %d", NTR_MAX_DIM);
    if (!nodes_out || !refs_out || !n_nodes_out || !n_refs_out || !root_out || !boundary_out) return ntr_fail(NTR_ERR_VALUE, "NULL output argument");
    if (n && (!lo || !hi)) return ntr_fail(NTR_ERR_VALUE, "item bounds are NULL");
    const int D = dim;
    for (size_t k = 0; k < (size_t)n * D; ++k)             // NaN, infinite or inverted bounds would poison the SAH sweep
        if (!(lo[k] <= hi[k]) || !(lo[k] > -3e38f) || !(hi[k] < 3e38f))
            return ntr_fail(NTR_ERR_VALUE, "item %zu: bounds are NaN, infinite or inverted", k / D);
    Builder b;
    b.D = D; b.lo = lo; b.hi = hi;
    b.item_first = item_first; b.s_lo = s_lo; b.s_hi = s_hi; b.s_rec = s_records;
    b.max_depth = max_depth > 0 ? std::min(max_depth, NTR_MAX_TREE_DEPTH - 2) : 25;
    b.split_threshold = split_threshold > 0 ? split_threshold : 2;
    b.c_trav = traversal_cost >= 0 ? traversal_cost : 1.0;
    b.c_isect = intersection_cost > 0 ? intersection_cost : 4.0;
    double blo[NTR_MAX_DIM], bhi[NTR_MAX_DIM];
    for (int i = 0; i < D; ++i) { blo[i] = 3e38; bhi[i] = -3e38; }
    for (uint32_t k = 0; k < n; ++k)
        for (int i = 0; i < D; ++i) {
            blo[i] = std::min(blo[i], (double)lo[(size_t)k * D + i]);
            bhi[i] = std::max(bhi[i], (double)hi[(size_t)k * D + i]);
        }
    for (int i = 0; i < D; ++i) {
        if (n == 0) { blo[i] = -1; bhi[i] = 1; }
        const double pad = 1e-5 * std::max(bhi[i] - blo[i], 1e-6);
        blo[i] -= pad; bhi[i] += pad;
        boundary_out[i] = (float)blo[i];
        boundary_out[D + i] = (float)bhi[i];
    }
    Builder::Tree tree;
    std::vector<uint32_t> idx(n);
    for (uint32_t k = 0; k < n; ++k) idx[k] = k;
    uint32_t root = NTR_NULL_NODE;
    try {
        // Parallel top: split the root sequentially a few levels down by building subtrees as tasks.  The simple and
        // robust way: build the whole tree in one task per top-level child (2 tasks), recursively up to `par_depth`.
        struct Par {
            const Builder &b;
            int par_depth;
            uint32_t run(Builder::Tree &t, std::vector<uint32_t> &ids, const double *nlo, const double *nhi, int depth) {
                if (depth >= par_depth || ids.size() < 50000) return b.build(t, ids, nlo, nhi, depth);
                // one split step here (re-using Builder::build on a depth-limited copy would duplicate work), so
                // emulate: build the two halves in parallel into private trees and splice them
                Builder one = b;
                one.max_depth = depth + 1;                   // forces children of this node to be leaves
                Builder::Tree top;
                std::vector<uint32_t> copy = ids;
                const uint32_t r = one.build(top, copy, nlo, nhi, depth);
                if (r == NTR_NULL_NODE || (top.nodes[r].meta & NTR_LEAF_FLAG)) return b.build(t, ids, nlo, nhi, depth);
                const ntr_node br = top.nodes[r];
                float split;
                memcpy(&split, &br.w1, 4);
                const int ax = (int)br.meta;
                auto child_ids = [&](uint32_t c) {
                    std::vector<uint32_t> v;
                    if (c != NTR_NULL_NODE) { const ntr_node &lf = top.nodes[c]; v.assign(top.refs.begin() + lf.w1, top.refs.begin() + lf.w1 + lf.w2); }
                    return v;
                };
                std::vector<uint32_t> li = child_ids(br.w2), ri = child_ids(br.w3);
                std::vector<uint32_t>().swap(ids);
                double l_hi[NTR_MAX_DIM], r_lo[NTR_MAX_DIM];
                memcpy(l_hi, nhi, sizeof(double) * b.D);
                memcpy(r_lo, nlo, sizeof(double) * b.D);
                l_hi[ax] = split; r_lo[ax] = split;
                Builder::Tree lt, rt;
                uint32_t lroot = NTR_NULL_NODE, rroot = NTR_NULL_NODE;
                auto fut = std::async(std::launch::async, [&] { lroot = run(lt, li, nlo, l_hi, depth + 1); });
                rroot = run(rt, ri, r_lo, nhi, depth + 1);
                fut.get();
                const uint32_t me = (uint32_t)t.nodes.size();
                t.nodes.push_back(ntr_node{(uint32_t)ax, br.w1, NTR_NULL_NODE, NTR_NULL_NODE});
                auto splice = [&](Builder::Tree &sub, uint32_t sroot) -> uint32_t {
                    if (sroot == NTR_NULL_NODE) return NTR_NULL_NODE;
                    const uint32_t nbase = (uint32_t)t.nodes.size(), rbase = (uint32_t)t.refs.size();
                    for (ntr_node nd : sub.nodes) {
                        if (nd.meta & NTR_LEAF_FLAG) nd.w1 += rbase;
                        else { if (nd.w2 != NTR_NULL_NODE) nd.w2 += nbase; if (nd.w3 != NTR_NULL_NODE) nd.w3 += nbase; }
                        t.nodes.push_back(nd);
                    }
                    t.refs.insert(t.refs.end(), sub.refs.begin(), sub.refs.end());
                    return sroot + nbase;
                };
                const uint32_t l = splice(lt, lroot), rr = splice(rt, rroot);
                t.nodes[me].w2 = l;
                t.nodes[me].w3 = rr;
                return me;
            }
        } par{b, 4};
        root = par.run(tree, idx, blo, bhi, 0);
    } catch (const std::bad_alloc &) {
        return ntr_fail(NTR_ERR_MEMORY, "out of memory building the k-d tree");
    } catch (const std::system_error &e) {
        return ntr_fail(NTR_ERR_RUNTIME, "k-d tree builder: %s", e.what());
    }
    *n_nodes_out = (uint32_t)tree.nodes.size();
    *n_refs_out = (uint32_t)tree.refs.size();
    *root_out = root;
    *nodes_out = (ntr_node *)malloc(std::max<size_t>(1, tree.nodes.size()) * sizeof(ntr_node));
    *refs_out = (uint32_t *)malloc(std::max<size_t>(1, tree.refs.size()) * sizeof(uint32_t));
    if (!*nodes_out || !*refs_out) { free(*nodes_out); free(*refs_out); return ntr_fail(NTR_ERR_MEMORY, "out of memory"); }
    if (!tree.nodes.empty()) memcpy(*nodes_out, tree.nodes.data(), tree.nodes.size() * sizeof(ntr_node));
    if (!tree.refs.empty()) memcpy(*refs_out, tree.refs.data(), tree.refs.size() * sizeof(uint32_t));
    return NTR_OK;
}

}  // extern "C"
