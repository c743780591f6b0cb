/* ntr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain, scalar, recursive C restatement of the reference renderer's per-pixel path
 * (Rouslan/NTracer, /root/reference).  It exists so that the CUDA path can be checked on the GPU box
 * (where /root/reference does not exist) and it is itself pinned against golden vectors produced by
 * the real reference (tests/golden/make_fixtures.py -> tests/golden/*.npz, tests/test_oracle_golden.py)
 * and against the reference's own known-answer test test_kdtree (lib/ntracer/tests/test.py:302-363).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's shared
 * object.  ntracer_b200/ never does: the product has no CPU path.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/src).
 * The structure deliberately mirrors the reference (recursive traversal, heap lists); it shares no
 * code with ntracer_b200/csrc (stack machine, fixed-size lists) -- only the POD scene description.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "../include/ntracer_b200.h"

#define MAXD NTR_MAX_DIM
#define NONE_REF 0xFFFFFFFFu

static const float ROUNDING_FUZZ = FLT_EPSILON * 10;        /* tracer.hpp:25 */
static const float LIGHT_THRESHOLD = 1.0f / 512;            /* tracer.hpp:31 */
#define QUICK_LIST_PREALLOC 10                               /* tracer.hpp:26 */
#define ALL_HITS_LIST_PREALLOC 20                            /* tracer.hpp:27 */

typedef struct { float o[MAXD], d[MAXD]; } oray;             /* ray<Store>, tracer.hpp:47-58 */
typedef struct { uint32_t ref; int lane; } otarget;          /* intersection_target<Store,true>, tracer.hpp:744-763 */
typedef struct { float dist; otarget target; oray normal; } ohit;   /* ray_intersection, tracer.hpp:765-779 */

/* quick_list<ray_intersection> (tracer.hpp:663-729).  The reference's check_capacity copies
 * alloc_size BYTES instead of elements (tracer.hpp:670-680): once a list grows past its
 * pre-allocated size its contents are undefined.  The oracle keeps a correct list and raises
 * `undefined` so tests can exclude those rays. */
typedef struct { ohit *v; size_t n, cap; } ohits;
typedef struct { uint32_t *v; size_t n, cap; } omail;        /* prim_list, tracer.hpp:782 */

typedef struct {
    const ntr_scene_desc *s;
    int D;
    size_t sstride, solstride;
    const float *cam_o, *right, *up, *fwd;
    float half_w, half_h, fovI;
    ntr_counters cnt;
    int undefined;      /* bit 0: a transparent-hit list, bit 1: the mailbox outgrew its preallocation during this pixel;
                         * bit 2: ill-conditioned -- an opaque hit was shaded at a point that is not on its own surface
                         * (see ray_color) */
} octx;

static void hits_add(octx *c, ohits *l, const ohit *h) {
    if (l->n >= QUICK_LIST_PREALLOC) c->undefined |= 1;      /* transparent-hit list outgrew its 10 slots */
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 16;
        l->v = (ohit *)realloc(l->v, l->cap * sizeof(ohit));
    }
    l->v[l->n++] = *h;
}
static void hits_remove_at(ohits *l, size_t i) {             /* quick_list::remove_at, tracer.hpp:723-728 */
    --l->n;
    if (i != l->n) l->v[i] = l->v[l->n];
}
static void mail_add(octx *c, omail *l, uint32_t ref) {
    if (l->n >= ALL_HITS_LIST_PREALLOC) c->undefined |= 2;   /* mailbox outgrew its 20 slots */
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 32;
        l->v = (uint32_t *)realloc(l->v, l->cap * sizeof(uint32_t));
    }
    l->v[l->n++] = ref;
}
static int mail_has(const omail *l, uint32_t ref) {          /* has(), tracer.hpp:832-834 */
    for (size_t i = 0; i < l->n; ++i) if (l->v[i] == ref) return 1;
    return 0;
}
/* trim_intersections, tracer.hpp:784-789 */
static void trim_intersections(ohits *l, float dist, size_t from) {
    while (from < l->n) {
        if (l->v[from].dist >= dist) hits_remove_at(l, from);
        else ++from;
    }
}
/* quick_list::sort_and_unique, tracer.hpp:714-721 (std::sort by dist; std::unique on equal target).
 * Insertion sort: stable, so equal distances keep insertion order (std::sort leaves it unspecified). */
static void sort_and_unique(ohits *l) {
    for (size_t i = 1; i < l->n; ++i) {
        ohit k = l->v[i];
        size_t j = i;
        while (j > 0 && k.dist < l->v[j - 1].dist) { l->v[j] = l->v[j - 1]; --j; }
        l->v[j] = k;
    }
    if (l->n == 0) return;
    size_t w = 0;
    for (size_t i = 1; i < l->n; ++i) {
        if (!(l->v[i].target.ref == l->v[w].target.ref && l->v[i].target.lane == l->v[w].target.lane))
            l->v[++w] = l->v[i];
    }
    l->n = w + 1;
}

static float dotD(int D, const float *a, const float *b) {
    float s = 0;
    for (int i = 0; i < D; ++i) s += a[i] * b[i];
    return s;
}
static void unitD(int D, const float *a, float *out) {       /* vector::unit, geometry.hpp:257-261 */
    float len = sqrtf(dotD(D, a, a));
    for (int i = 0; i < D; ++i) out[i] = a[i] / len;
}

/* ---- materials ------------------------------------------------------------------------------- */
typedef struct { float c[3], specular[3], opacity, reflectivity, specular_intensity, specular_exp; } omat;

static const omat *target_mat(const octx *c, otarget t) {    /* intersection_target::mat, tracer.hpp:752-762 */
    uint32_t kind = t.ref >> 30, idx = t.ref & 0x3FFFFFFFu;
    int32_t m;
    if (kind == NTR_REF_BATCH) m = c->s->simplex_mat[idx + (uint32_t)t.lane];
    else if (kind == NTR_REF_SIMPLEX) m = c->s->simplex_mat[idx];
    else m = c->s->solid_mat[idx];
    return (const omat *)(c->s->materials + (size_t)m * 10);
}
static int target_opaque(const octx *c, otarget t) { return target_mat(c, t)->opacity >= 1; }   /* tracer.hpp:187-189,211-213 */

/* ---- primitives ------------------------------------------------------------------------------ */
/* triangle::intersects, tracer.hpp:411-440 */
static float triangle_intersects(octx *c, uint32_t idx, const oray *target, oray *normal, float cutoff) {
    const int D = c->D;
    const float *fn = c->s->simplex + (size_t)idx * c->sstride;
    const float d = fn[D];
    const float *p1 = fn + D + 1;
    const float *edges = fn + 2 * D + 1;
    c->cnt.simplex_tests++;

    float denom = dotD(D, fn, target->d);
    if (!denom) return 0;
    float t = -(dotD(D, fn, target->o) + d) / denom;
    if (t <= 0 || t >= cutoff) return 0;

    float P[MAXD], pside[MAXD];
    for (int i = 0; i < D; ++i) { P[i] = target->o[i] + t * target->d[i]; pside[i] = p1[i] - P[i]; }

    float tot_area = 0;
    for (int i = 0; i < D - 1; ++i) {
        float area = dotD(D, edges + i * D, pside);
        if (area < -ROUNDING_FUZZ || area > (1 + ROUNDING_FUZZ)) return 0;
        tot_area += area;
    }
    if (tot_area <= (1 + ROUNDING_FUZZ)) {
        memcpy(normal->o, P, sizeof(float) * D);
        unitD(D, fn, normal->d);
        if (denom > 0) for (int i = 0; i < D; ++i) normal->d[i] = -normal->d[i];
        return t;
    }
    return 0;
}

/* triangle_batch::intersects, tracer.hpp:551-599 (one SIMD lane per simplex record) */
static float batch_intersects(octx *c, uint32_t first, const oray *target, oray *normal, int *index, float cutoff) {
    const int D = c->D, B = c->s->batch_size;
    float t[64], denoms[64];
    for (int l = 0; l < B; ++l) {
        const float *fn = c->s->simplex + (size_t)(first + l) * c->sstride;
        const float d = fn[D];
        const float *p1 = fn + D + 1;
        const float *edges = fn + 2 * D + 1;
        c->cnt.simplex_tests++;
        float denom = dotD(D, fn, target->d);
        int mask = denom != 0;
        float tl = -(dotD(D, fn, target->o) + d) / denom;
        mask = mask && tl >= 0;
        float pside[MAXD];
        for (int i = 0; i < D; ++i) pside[i] = p1[i] - (target->o[i] + tl * target->d[i]);
        float tot_area = 0;
        for (int i = 0; i < D - 1; ++i) {
            float area = dotD(D, edges + i * D, pside);
            mask = mask && area >= -ROUNDING_FUZZ;
            tot_area += area;
        }
        mask = mask && tot_area <= (1 + ROUNDING_FUZZ);
        t[l] = mask ? tl : 0;
        denoms[l] = denom;
    }
    float min_t = cutoff;
    int r_index = -1;
    for (int i = 0; i < B; ++i) {
        if (i != *index && t[i] && t[i] < min_t) { min_t = t[i]; r_index = i; }
    }
    if (r_index == -1) return 0;
    *index = r_index;
    const float *fn = c->s->simplex + (size_t)(first + r_index) * c->sstride;
    for (int i = 0; i < D; ++i) normal->o[i] = target->o[i] + min_t * target->d[i];
    unitD(D, fn, normal->d);
    if (denoms[r_index] > 0) for (int i = 0; i < D; ++i) normal->d[i] = -normal->d[i];
    return min_t;
}

/* hypercube_intersects, tracer.hpp:126-152.  Note: writes into normal->o even when it misses. */
static float hypercube_intersects(int D, const oray *target, oray *normal, float cutoff) {
    for (int i = 0; i < D; ++i) {
        if (target->d[i]) {
            normal->o[i] = target->d[i] < 0 ? 1.0f : -1.0f;
            float dist = (normal->o[i] - target->o[i]) / target->d[i];
            if (dist > 0) {
                int miss = 0;
                for (int j = 0; j < D; ++j) {
                    if (i != j) {
                        normal->o[j] = target->d[j] * dist + target->o[j];
                        if (fabsf(normal->o[j]) > (1 + ROUNDING_FUZZ)) { miss = 1; break; }
                    }
                }
                if (!miss) {
                    if (dist >= cutoff) return 0;
                    for (int j = 0; j < D; ++j) normal->d[j] = 0;
                    normal->d[i] = normal->o[i];
                    return dist;
                }
            }
        }
    }
    return 0;
}

/* hypersphere_intersects, tracer.hpp:154-173 */
static float hypersphere_intersects(int D, const oray *target, oray *normal, float cutoff) {
    float a = dotD(D, target->d, target->d);
    float b = 2 * dotD(D, target->d, target->o);
    float cc = dotD(D, target->o, target->o) - 1;
    float discriminant = b * b - 4 * a * cc;
    if (discriminant < 0) return 0;
    float dist = (-b - sqrtf(discriminant)) / (2 * a);
    if (dist <= 0 || dist >= cutoff) return 0;
    for (int i = 0; i < D; ++i) normal->d[i] = normal->o[i] = target->o[i] + target->d[i] * dist;
    return dist;
}

/* solid::intersects, tracer.hpp:251-276 */
static float solid_intersects(octx *c, uint32_t idx, const oray *target, oray *normal, float cutoff) {
    const int D = c->D;
    const float *rec = c->s->solids + (size_t)idx * c->solstride;
    const int type = (int)rec[0];
    const float *orient = rec + 1, *inv = rec + 1 + D * D, *pos = rec + 1 + 2 * D * D;
    c->cnt.solid_tests++;
    oray tr;
    for (int i = 0; i < D; ++i) {
        tr.o[i] = dotD(D, inv + i * D, target->o) - pos[i];
        tr.d[i] = dotD(D, inv + i * D, target->d);
    }
    float dist;
    if (type == NTR_SOLID_CUBE) dist = hypercube_intersects(D, &tr, normal, cutoff);
    else dist = hypersphere_intersects(D, &tr, normal, cutoff);
    if (!dist) return 0;
    float tmp[MAXD], no[MAXD], nd[MAXD];
    for (int i = 0; i < D; ++i) tmp[i] = normal->o[i] + pos[i];
    for (int i = 0; i < D; ++i) { no[i] = dotD(D, orient + i * D, tmp); nd[i] = dotD(D, orient + i * D, normal->d); }
    memcpy(normal->o, no, sizeof(float) * D);
    memcpy(normal->d, nd, sizeof(float) * D);
    return dist;
}

/* primitive::intersects dispatch, tracer.hpp:508-516 */
static float prim_intersects(octx *c, uint32_t ref, const oray *target, oray *normal, float cutoff) {
    if ((ref >> 30) == NTR_REF_SIMPLEX) return triangle_intersects(c, ref & 0x3FFFFFFFu, target, normal, cutoff);
    return solid_intersects(c, ref & 0x3FFFFFFFu, target, normal, cutoff);
}

/* ---- k-d tree -------------------------------------------------------------------------------- */
static float node_split(const ntr_node *n) { float f; memcpy(&f, &n->w1, 4); return f; }

/* kd_leaf<Store,true>::intersects, tracer.hpp:977-1086 -- the batched leaf, which is what every SIMD
 * build of the reference instantiates (the scalar variant :858-913 differs only in that it does not
 * re-test the primitive that produced the first opaque hit).  Mirrored literally, including:
 *   - before the first opaque hit the tests write straight into o_hit.normal (:1001,1020);
 *   - `goto hit` leaves `i` un-incremented and skips checked.add, so the hitting item is tested a
 *     second time with cutoff == its own distance (and misses);
 *   - the final trim uses the result of the LAST test, not o_hit.dist (:1084). */
static int leaf_intersects(octx *c, const ntr_node *leaf, const oray *target, otarget skip,
                           ohit *o_hit, ohits *t_hits, omail *checked) {
    const uint32_t *items = c->s->leaf_refs + leaf->w1;
    const uint32_t size = leaf->w2;
    size_t h_start = t_hits->n;
    float dist = 0;
    uint32_t i = 0;
    oray new_normal;

    for (; i < size; ++i) {
        uint32_t item = items[i];
        if ((item >> 30) == NTR_REF_BATCH) {
            if (!mail_has(checked, item)) {
                int index = skip.ref == item ? skip.lane : -1;
                dist = batch_intersects(c, item & 0x3FFFFFFFu, target, &o_hit->normal, &index, o_hit->dist);
                if (dist) {
                    otarget tg = { item, index };
                    if (target_opaque(c, tg)) {
                        o_hit->dist = dist;
                        o_hit->target = tg;
                        goto hit;
                    }
                    ohit h = { dist, tg, o_hit->normal };
                    hits_add(c, t_hits, &h);
                }
                mail_add(c, checked, item);
            }
        } else if (item != skip.ref && !mail_has(checked, item)) {
            dist = prim_intersects(c, item, target, &o_hit->normal, o_hit->dist);
            if (dist) {
                otarget tg = { item, -1 };
                if (target_opaque(c, tg)) {
                    o_hit->dist = dist;
                    o_hit->target = tg;
                    goto hit;
                }
                ohit h = { dist, tg, o_hit->normal };
                hits_add(c, t_hits, &h);
            }
            mail_add(c, checked, item);
        }
    }
    return 0;

hit:
    memset(&new_normal, 0, sizeof new_normal);
    for (; i < size; ++i) {
        uint32_t item = items[i];
        if ((item >> 30) == NTR_REF_BATCH) {
            if (!mail_has(checked, item)) {
                int index = skip.ref == item ? skip.lane : -1;
                dist = batch_intersects(c, item & 0x3FFFFFFFu, target, &new_normal, &index, o_hit->dist);
                if (dist) {
                    otarget tg = { item, index };
                    if (target_opaque(c, tg)) {
                        o_hit->dist = dist;
                        o_hit->normal = new_normal;
                        o_hit->target = tg;
                    } else {
                        ohit h = { dist, tg, new_normal };
                        hits_add(c, t_hits, &h);
                    }
                }
                mail_add(c, checked, item);
            }
        } else if (item != skip.ref && !mail_has(checked, item)) {
            dist = prim_intersects(c, item, target, &new_normal, o_hit->dist);
            if (dist) {
                otarget tg = { item, -1 };
                if (target_opaque(c, tg)) {
                    o_hit->dist = dist;
                    o_hit->normal = new_normal;
                    o_hit->target = tg;
                } else {
                    ohit h = { dist, tg, new_normal };
                    hits_add(c, t_hits, &h);
                }
            }
            mail_add(c, checked, item);
        }
    }
    trim_intersections(t_hits, dist, h_start);
    return 1;
}

typedef struct {
    octx *c; const oray *target; float invdir[MAXD]; otarget skip; ohit *o_hit; ohits *t_hits; omail checked;
} kd_isect;

/* kd_node_intersection::operator(), tracer.hpp:1179-1243 */
static int kd_intersect(kd_isect *k, uint32_t node, float t_near, float t_far) {
    octx *c = k->c;
    const oray *target = k->target;
    while (node != NTR_NULL_NODE) {
        const ntr_node *n = c->s->nodes + node;
        if (n->meta & NTR_LEAF_FLAG)
            return leaf_intersects(c, n, target, k->skip, k->o_hit, k->t_hits, &k->checked);
        c->cnt.node_steps++;
        const uint32_t axis = n->meta;
        const float split = node_split(n);
        const uint32_t left = n->w2, right = n->w3;
        if (target->d[axis]) {
            if (target->o[axis] == split) {
                node = target->d[axis] > 0 ? right : left;
                continue;
            }
            float t = (split - target->o[axis]) * k->invdir[axis];
            uint32_t n_near = target->o[axis] > split ? right : left;
            uint32_t n_far = target->o[axis] > split ? left : right;
            if (t < 0 || t > t_far) { node = n_near; continue; }
            if (t < t_near) { node = n_far; continue; }
            if (n_near != NTR_NULL_NODE) {
                size_t h_start = k->t_hits->n;
                int hit = kd_intersect(k, n_near, t_near, t);
                if ((hit && k->o_hit->dist <= t) || n_far == NTR_NULL_NODE) return hit;
                if (hit) {
                    if (kd_intersect(k, n_far, t, t_far))
                        trim_intersections(k->t_hits, k->o_hit->dist, h_start);
                    return 1;
                }
            }
            node = n_far;
            t_near = t;
            continue;
        }
        node = target->o[axis] >= split ? right : left;
    }
    return 0;
}

/* intersects(), tracer.hpp:1245-1256 */
static int tree_intersects(octx *c, uint32_t root, const oray *target, otarget skip, ohit *o_hit, ohits *t_hits,
                           float t_near, float t_far) {
    kd_isect k;
    k.c = c; k.target = target; k.skip = skip; k.o_hit = o_hit; k.t_hits = t_hits;
    k.checked.v = NULL; k.checked.n = k.checked.cap = 0;
    for (int i = 0; i < c->D; ++i) k.invdir[i] = 1 / target->d[i];      /* tracer.hpp:1174 */
    int r = kd_intersect(&k, root, t_near, t_far);
    free(k.checked.v);
    return r;
}

/* kd_leaf<Store,true>::occludes, tracer.hpp:1088-1124 (scalar :915-939) */
static int leaf_occludes(octx *c, const ntr_node *leaf, const oray *target, float ldistance, otarget skip, ohits *hits) {
    const uint32_t *items = c->s->leaf_refs + leaf->w1;
    const uint32_t size = leaf->w2;
    oray normal;
    memset(&normal, 0, sizeof normal);
    for (uint32_t i = 0; i < size; ++i) {
        uint32_t item = items[i];
        if ((item >> 30) == NTR_REF_BATCH) {
            int index = skip.ref == item ? skip.lane : -1;
            float dist = batch_intersects(c, item & 0x3FFFFFFFu, target, &normal, &index, ldistance);
            if (dist) {
                otarget tg = { item, index };
                if (target_opaque(c, tg)) return 1;
                ohit h = { dist, tg, normal };
                hits_add(c, hits, &h);
            }
        } else if (item != skip.ref) {
            float dist = prim_intersects(c, item, target, &normal, ldistance);
            if (dist) {
                otarget tg = { item, -1 };
                if (target_opaque(c, tg)) return 1;
                ohit h = { dist, tg, normal };
                hits_add(c, hits, &h);
            }
        }
    }
    return 0;
}

/* _occludes, tracer.hpp:1258-1307 */
static int kd_occludes(octx *c, uint32_t node, const oray *target, const float *invdir, float ldistance,
                       otarget skip, ohits *hits, float t_near, float t_far) {
    while (node != NTR_NULL_NODE) {
        const ntr_node *n = c->s->nodes + node;
        if (n->meta & NTR_LEAF_FLAG) return leaf_occludes(c, n, target, ldistance, skip, hits);
        c->cnt.node_steps++;
        const uint32_t axis = n->meta;
        const float split = node_split(n);
        const uint32_t left = n->w2, right = n->w3;
        if (target->d[axis]) {
            if (target->o[axis] == split) {
                node = target->d[axis] > 0 ? right : left;
                continue;
            }
            float t = (split - target->o[axis]) * invdir[axis];
            uint32_t n_near = left, n_far = right;
            if (target->o[axis] > split) { n_near = right; n_far = left; }
            if (t < 0 || t > t_far) { node = n_near; continue; }
            if (t < t_near) { node = n_far; continue; }
            if (n_near != NTR_NULL_NODE) {
                if (n_far == NTR_NULL_NODE) { t_far = t; node = n_near; continue; }
                if (kd_occludes(c, n_near, target, invdir, ldistance, skip, hits, t_near, t)) return 1;
            }
            if (t < ldistance) return 0;            /* tracer.hpp:1298 */
            t_near = t;
            node = n_far;
            continue;
        }
        node = target->o[axis] >= split ? right : left;
    }
    return 0;
}

/* occludes(), tracer.hpp:1309-1311 */
static int tree_occludes(octx *c, uint32_t root, const oray *target, float ldistance, otarget skip, ohits *hits,
                         float t_near, float t_far) {
    float invdir[MAXD];
    for (int i = 0; i < c->D; ++i) invdir[i] = 1 / target->d[i];
    return kd_occludes(c, root, target, invdir, ldistance, skip, hits, t_near, t_far);
}

/* ---- composite_scene ------------------------------------------------------------------------- */
typedef struct { float r, g, b; } ocolor;
static ocolor col(float r, float g, float b) { ocolor c = { r, g, b }; return c; }
static ocolor cmulf(ocolor a, float f) { return col(a.r * f, a.g * f, a.b * f); }
static ocolor cmul(ocolor a, ocolor b) { return col(a.r * b.r, a.g * b.g, a.b * b.b); }
static ocolor cadd(ocolor a, ocolor b) { return col(a.r + b.r, a.g + b.g, a.b + b.b); }
static ocolor col3(const float *p) { return col(p[0], p[1], p[2]); }

static ocolor ray_color(octx *c, const oray *target, int depth, otarget source);

/* composite_scene::light_reaches, tracer.hpp:1750-1766 */
static int light_reaches(octx *c, const oray *target, float ldistance, otarget skip, ocolor *filtered) {
    ohits th = { NULL, 0, 0 };
    c->cnt.shadow_rays++;
    if (tree_occludes(c, c->s->root, target, ldistance, skip, &th, 0, FLT_MAX)) { free(th.v); return 0; }
    if (th.n) {
        sort_and_unique(&th);
        for (size_t i = th.n; i-- > 0;) {
            float f = 1 - target_mat(c, th.v[i].target)->opacity;
            *filtered = cmulf(*filtered, f);
        }
    }
    free(th.v);
    return 1;
}

/* append_specular, tracer.hpp:1701-1707 */
static void append_specular(int D, ocolor *cs, float *a, const omat *m, ocolor light_c, const float *target,
                            const float *normal, const float *light_dir) {
    float h[MAXD], hu[MAXD];
    for (int i = 0; i < D; ++i) h[i] = light_dir[i] - target[i];
    unitD(D, h, hu);
    float base = powf(dotD(D, normal, hu), m->specular_exp) * m->specular_intensity;
    *cs = cadd(*cs, cmulf(cmul(col3(m->specular), light_c), base * (1 - *a)));
    *a += base * (1 - *a);
    *cs = cmulf(*cs, *a);
}

/* composite_scene::base_color, tracer.hpp:1768-1854 */
static ocolor base_color(octx *c, const oray *target, const oray *normal, otarget source, int depth) {
    const int D = c->D;
    const ntr_scene_desc *s = c->s;
    const omat *m = target_mat(c, source);
    ocolor light = col(0, 0, 0), specular = col(0, 0, 0);
    float spec_a = 0;
    c->cnt.shaded_hits++;

    for (uint32_t li = 0; li < s->n_point_lights; ++li) {
        const float *pl = s->point_lights + (size_t)li * (D + 3);
        ocolor plc = col3(pl + D);
        float lv[MAXD];
        for (int i = 0; i < D; ++i) lv[i] = normal->o[i] - pl[i];
        float dist = sqrtf(dotD(D, lv, lv));
        for (int i = 0; i < D; ++i) lv[i] /= dist;
        float sine = dotD(D, normal->d, lv);
        if (sine > 0) {
            float strength = (float)(1 / pow((double)dist, (double)(D - 1)));       /* tracer.hpp:1686-1688 */
            if (s->shadows) {
                if (fmaxf(plc.r, fmaxf(plc.g, plc.b)) * strength * sine > LIGHT_THRESHOLD) {
                    ocolor filtered = plc;
                    oray sr;
                    memcpy(sr.o, normal->o, sizeof(float) * D);
                    memcpy(sr.d, lv, sizeof(float) * D);
                    if (light_reaches(c, &sr, dist, source, &filtered)) {
                        filtered = cmulf(filtered, strength);
                        light = cadd(light, cmulf(filtered, sine));
                        if (m->specular_intensity) append_specular(D, &specular, &spec_a, m, filtered, target->d, normal->d, lv);
                    }
                }
            } else {
                light = cadd(light, cmulf(cmulf(plc, strength), sine));
            }
        }
    }
    for (uint32_t li = 0; li < s->n_global_lights; ++li) {
        const float *gl = s->global_lights + (size_t)li * (D + 3);
        ocolor glc = col3(gl + D);
        float sine = -dotD(D, normal->d, gl);
        if (sine > 0) {
            if (s->shadows) {
                ocolor filtered = glc;
                oray sr;
                memcpy(sr.o, normal->o, sizeof(float) * D);
                for (int i = 0; i < D; ++i) sr.d[i] = -gl[i];
                if (light_reaches(c, &sr, FLT_MAX, source, &filtered)) {
                    light = cadd(light, cmulf(filtered, sine));
                    if (m->specular_intensity) append_specular(D, &specular, &spec_a, m, filtered, target->d, normal->d, sr.d);
                }
            } else {
                light = cadd(light, cmulf(glc, sine));
            }
        }
    }

    float sine = -dotD(D, target->d, normal->d);
    if (s->camera_light && sine > 0) {
        light = cadd(light, col(sine, sine, sine));
        if (m->specular_intensity) {
            float base = powf(sine, m->specular_exp) * m->specular_intensity;
            specular = cadd(specular, cmulf(col3(m->specular), base * (1 - spec_a)));
            spec_a += base * (1 - spec_a);
            specular = cmulf(specular, spec_a);
        }
    }

    ocolor r = cadd(col3(s->ambient), cmul(col3(m->c), light));

    if (m->reflectivity && depth < s->max_reflect_depth) {
        oray rr;
        memcpy(rr.o, normal->o, sizeof(float) * D);
        for (int i = 0; i < D; ++i) rr.d[i] = target->d[i] - normal->d[i] * (-2 * sine);
        c->cnt.reflection_rays++;
        ocolor rc = ray_color(c, &rr, depth + 1, source);
        r = cadd(cmulf(cmul(col3(m->c), rc), m->reflectivity), cmulf(r, 1 - m->reflectivity));
    }
    return cadd(specular, cmulf(r, 1 - spec_a));
}

/* composite_scene::aabb_distance, tracer.hpp:1892-1918 */
static float aabb_distance(const octx *c, const oray *target) {
    const int D = c->D;
    const float *start = c->s->boundary, *end = c->s->boundary + D;
    for (int i = 0; i < D; ++i) {
        if (target->d[i]) {
            float o = target->d[i] > 0 ? start[i] : end[i];
            float dist = (o - target->o[i]) / target->d[i];
            int skip = i;
            if (dist < 0) { dist = 0; skip = -1; }
            int miss = 0;
            for (int j = 0; j < D; ++j) {
                if (j != skip) {
                    o = target->d[j] * dist + target->o[j];
                    if (o >= end[j] || o <= start[j]) { miss = 1; break; }
                }
            }
            if (!miss) return dist;
        }
    }
    return -1;
}

/* composite_scene::ray_color, tracer.hpp:1856-1883 */
static ocolor ray_color(octx *c, const oray *target, int depth, otarget source) {
    const ntr_scene_desc *s = c->s;
    ohit hit;
    memset(&hit, 0, sizeof hit);
    ohits th = { NULL, 0, 0 };
    ocolor r;

    float dist = aabb_distance(c, target);
    hit.dist = FLT_MAX;
    if (dist >= 0 && tree_intersects(c, s->root, target, source, &hit, &th, dist, FLT_MAX)) {
        /* kd_leaf::intersects lets every test made before the first opaque hit of a leaf write straight into
         * o_hit.normal (tracer.hpp:1001,1020), so a transparent hit -- or a hypercube test that misses after writing
         * some coordinates (:131-139) -- in a leaf visited later replaces the shading point of an opaque hit found
         * earlier.  The secondary rays of such a pixel start ON that other surface while `source` names the opaque
         * primitive, so whether they re-hit the surface they start on at t ~ 0 is decided by the last bit of t: the
         * reference's own answer depends on its build (-ffast-math, FMA contraction).  Restated as it is, and
         * flagged so that parity tests can tell these pixels apart. */
        {
            oray clean;
            memset(&clean, 0, sizeof clean);
            int same = 1;
            if ((hit.target.ref >> 30) == NTR_REF_SOLID) {
                ntr_counters keep = c->cnt;
                solid_intersects(c, hit.target.ref & 0x3FFFFFFFu, target, &clean, FLT_MAX);
                c->cnt = keep;
                for (int i = 0; i < c->D; ++i) same = same && clean.o[i] == hit.normal.o[i];
            } else {
                for (int i = 0; i < c->D; ++i) same = same && (target->o[i] + hit.dist * target->d[i]) == hit.normal.o[i];
            }
            if (!same) c->undefined |= 4;
        }
        r = base_color(c, target, &hit.normal, hit.target, depth);
    } else {
        float intensity = target->d[s->bg_gradient_axis];
        r = intensity >= 0 ? cadd(cmulf(col3(s->bg1), intensity), cmulf(col3(s->bg2), 1 - intensity))
                           : cadd(cmulf(col3(s->bg3), -intensity), cmulf(col3(s->bg2), 1 + intensity));
    }
    if (th.n) {
        sort_and_unique(&th);
        for (size_t i = th.n; i-- > 0;) {
            const omat *m = target_mat(c, th.v[i].target);
            ocolor base = base_color(c, target, &th.v[i].normal, th.v[i].target, depth);
            r = cadd(cmulf(base, m->opacity), cmulf(r, 1 - m->opacity));
        }
    }
    free(th.v);
    return r;
}

/* flat_origin_ray_source, tracer.hpp:60-76 */
static void set_view(octx *c, int w, int h) {
    c->half_w = (float)w / 2.0f;
    c->half_h = (float)h / 2.0f;
    c->fovI = tanf(c->s->fov / 2) / c->half_w;
}
static void primary_dir(const octx *c, float x, float y, float *out) {
    float v[MAXD];
    float fx = c->fovI * (x - c->half_w), fy = c->fovI * (y - c->half_h);
    for (int i = 0; i < c->D; ++i) v[i] = c->fwd[i] + c->right[i] * fx - c->up[i] * fy;
    unitD(c->D, v, out);
}

/* box_scene::calculate_color, tracer.hpp:101-114 ; composite_scene::calculate_color, :1885-1890 */
static ocolor calculate_color(octx *c, int x, int y) {
    const int D = c->D;
    oray view;
    memcpy(view.o, c->cam_o, sizeof(float) * D);
    primary_dir(c, (float)x, (float)y, view.d);
    c->cnt.primary_rays++;
    if (c->s->kind == NTR_SCENE_BOX) {
        oray normal;
        memset(&normal, 0, sizeof normal);
        if (hypercube_intersects(D, &view, &normal, FLT_MAX)) {
            float sine = dotD(D, view.d, normal.d);
            return cmulf(col(1, 0.5f, 0.5f), sine <= 0 ? -sine : 0.0f);
        }
        float intensity = view.d[0];
        return intensity > 0 ? col(intensity, intensity, intensity) : col(0, -intensity, -intensity);
    }
    otarget none = { NONE_REF, 0 };
    return ray_color(c, &view, 0, none);
}

static void ctx_init(octx *c, const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes) {
    memset(c, 0, sizeof *c);
    c->s = s;
    c->D = s->dim;
    c->sstride = (size_t)(s->dim + 1) * s->dim + 1;
    c->solstride = 1 + 2 * (size_t)s->dim * s->dim + s->dim;
    c->cam_o = cam_origin;
    if (cam_axes) { c->right = cam_axes; c->up = cam_axes + s->dim; c->fwd = cam_axes + 2 * s->dim; }
}

static void cnt_add(ntr_counters *a, const ntr_counters *b) {
    a->primary_rays += b->primary_rays; a->reflection_rays += b->reflection_rays; a->shadow_rays += b->shadow_rays;
    a->node_steps += b->node_steps; a->simplex_tests += b->simplex_tests; a->solid_tests += b->solid_tests;
    a->shaded_hits += b->shaded_hits;
}

static int flat_id(const ntr_scene_desc *s, otarget t) {
    uint32_t kind = t.ref >> 30, idx = t.ref & 0x3FFFFFFFu;
    if (kind == NTR_REF_BATCH) return (int)(idx + (uint32_t)t.lane);
    if (kind == NTR_REF_SIMPLEX) return (int)idx;
    return (int)(s->n_simplex + idx);
}

/* ---- exported entry points (ctypes) ----------------------------------------------------------- */
#define ORACLE_API __attribute__((visibility("default")))

/* Float image; rows [y0, y1) x columns [x0, x1) of a w x h view, written at their frame position.
 * undefined_mask (optional, w*h bytes) is set to 1 for pixels where the reference itself is undefined. */
ORACLE_API int oracle_render_float_window(const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes,
                                          int w, int h, int x0, int y0, int x1, int y1,
                                          float *rgb, uint8_t *undefined_mask, ntr_counters *counters) {
    ntr_counters total;
    memset(&total, 0, sizeof total);
#pragma omp parallel
    {
        octx c;
        ctx_init(&c, s, cam_origin, cam_axes);
        set_view(&c, w, h);
#pragma omp for schedule(dynamic, 1)
        for (int y = y0; y < y1; ++y) {
            for (int x = x0; x < x1; ++x) {
                c.undefined = 0;
                ocolor r = calculate_color(&c, x, y);
                float *p = rgb + ((size_t)y * w + x) * 3;
                p[0] = r.r; p[1] = r.g; p[2] = r.b;
                if (undefined_mask) undefined_mask[(size_t)y * w + x] = (uint8_t)c.undefined;
            }
        }
#pragma omp critical
        cnt_add(&total, &c.cnt);
    }
    if (counters) *counters = total;
    return 0;
}

ORACLE_API int oracle_render_float(const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes,
                                   int w, int h, float *rgb, uint8_t *undefined_mask, ntr_counters *counters) {
    return oracle_render_float_window(s, cam_origin, cam_axes, w, h, 0, 0, w, h, rgb, undefined_mask, counters);
}

ORACLE_API int oracle_calculate_color(const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes,
                                      int x, int y, int w, int h, float out[3]) {
    octx c;
    ctx_init(&c, s, cam_origin, cam_axes);
    set_view(&c, w, h);
    ocolor r = calculate_color(&c, x, y);
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
    return c.undefined;
}

/* process_pixel::operator(), render.cpp:406-465, for one pixel (Size == 1 path). */
static void pack_pixel(const ntr_image_format *f, const float *rgb, uint8_t *out) {
    uint64_t temp[2] = { 0, 0 };
    int b_offset = 0;
    for (int ci = 0; ci < f->n_channels; ++ci) {
        const ntr_channel *ch = &f->channels[ci];
        float v = ch->f_r * rgb[0] + ch->f_g * rgb[1] + ch->f_b * rgb[2] + ch->f_c;
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);            /* simd::clamp(...,0,1) */
        uint64_t ival;
        if (ch->tfloat) { uint32_t u; memcpy(&u, &v, 4); ival = u; }
        else ival = (uint64_t)lround(v * (double)(0xffffffffu >> (32 - ch->bit_size)));
        int o = b_offset / 64, rm = b_offset % 64;
        int sh = 64 - rm - ch->bit_size;
        temp[o] |= sh >= 0 ? ival << sh : ival >> -sh;
        if (rm + ch->bit_size > 64) temp[o + 1] = ival << (128 - rm - ch->bit_size);
        b_offset += ch->bit_size;
    }
    if (f->reversed) {
        for (int j = f->bytes_per_pixel - 1; j >= 0; --j) *out++ = (uint8_t)(temp[j / 8] >> ((7 - (j % 8)) * 8));
    } else {
        for (int j = 0; j < f->bytes_per_pixel; ++j) *out++ = (uint8_t)(temp[j / 8] >> ((7 - (j % 8)) * 8));
    }
}

/* worker_draw's addressing, render.cpp:482-490: row y at dst + y*pitch, pixel x at + x*bytes_per_pixel */
ORACLE_API int oracle_pack(const ntr_image_format *f, const float *rgb, uint8_t *dst) {
    for (int y = 0; y < f->height; ++y)
        for (int x = 0; x < f->width; ++x)
            pack_pixel(f, rgb + ((size_t)y * f->width + x) * 3, dst + (size_t)y * f->pitch + (size_t)x * f->bytes_per_pixel);
    return 0;
}

ORACLE_API int oracle_render_packed(const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes,
                                    const ntr_image_format *f, uint8_t *dst, ntr_counters *counters) {
    float *rgb = (float *)malloc(sizeof(float) * 3 * (size_t)f->width * f->height);
    if (!rgb) return -2;
    oracle_render_float(s, cam_origin, cam_axes, f->width, f->height, rgb, NULL, counters);
    oracle_pack(f, rgb, dst);
    free(rgb);
    return 0;
}

/* KDNode.intersects, ntracer_body.hpp:1412-1458, for n rays (origins/dirs n x D).
 * skip_ref may be NULL.  ids_out: flat id of the opaque hit or -1. */
ORACLE_API int oracle_trace_rays(const ntr_scene_desc *s, uint32_t n, const float *origins, const float *dirs,
                                 float t_near, float t_far, const uint32_t *skip_ref, const int32_t *skip_lane,
                                 int32_t *ids_out, float *dist_out, int32_t *n_transparent_out) {
#pragma omp parallel
    {
        octx c;
        ctx_init(&c, s, NULL, NULL);
#pragma omp for schedule(dynamic, 64)
        for (uint32_t i = 0; i < n; ++i) {
            oray ray;
            memcpy(ray.o, origins + (size_t)i * c.D, sizeof(float) * c.D);
            memcpy(ray.d, dirs + (size_t)i * c.D, sizeof(float) * c.D);
            otarget skip = { skip_ref ? skip_ref[i] : NONE_REF, skip_lane ? skip_lane[i] : -1 };
            ohit hit;
            memset(&hit, 0, sizeof hit);
            hit.dist = FLT_MAX;
            ohits th = { NULL, 0, 0 };
            int did = tree_intersects(&c, s->root, &ray, skip, &hit, &th, t_near, t_far);
            ids_out[i] = did ? flat_id(s, hit.target) : -1;
            if (dist_out) dist_out[i] = did ? hit.dist : 0;
            if (n_transparent_out) n_transparent_out[i] = (int32_t)th.n;
            free(th.v);
        }
    }
    return 0;
}

/* KDNode.occludes, ntracer_body.hpp:1460-1496 */
ORACLE_API int oracle_occludes_rays(const ntr_scene_desc *s, uint32_t n, const float *origins, const float *dirs,
                                    const float *distance, const uint32_t *skip_ref, const int32_t *skip_lane,
                                    int32_t *occluded_out, int32_t *n_transparent_out) {
#pragma omp parallel
    {
        octx c;
        ctx_init(&c, s, NULL, NULL);
#pragma omp for schedule(dynamic, 64)
        for (uint32_t i = 0; i < n; ++i) {
            oray ray;
            memcpy(ray.o, origins + (size_t)i * c.D, sizeof(float) * c.D);
            memcpy(ray.d, dirs + (size_t)i * c.D, sizeof(float) * c.D);
            otarget skip = { skip_ref ? skip_ref[i] : NONE_REF, skip_lane ? skip_lane[i] : -1 };
            ohits th = { NULL, 0, 0 };
            /* the Python binding's default t_near is lowest() (ntracer_body.hpp:1470) */
            int occ = tree_occludes(&c, s->root, &ray, distance ? distance[i] : FLT_MAX, skip, &th, -FLT_MAX, FLT_MAX);
            occluded_out[i] = occ;
            if (n_transparent_out) n_transparent_out[i] = occ ? 0 : (int32_t)th.n;
            free(th.v);
        }
    }
    return 0;
}

/* Primary-ray hit ids as the render path sees them: intersects(root, primary ray, {}, ..., aabb_distance, max)
 * (tracer.hpp:1861-1863). */
ORACLE_API int oracle_primary_hit_ids(const ntr_scene_desc *s, const float *cam_origin, const float *cam_axes,
                                      int w, int h, int32_t *ids_out, float *dist_out) {
#pragma omp parallel
    {
        octx c;
        ctx_init(&c, s, cam_origin, cam_axes);
        set_view(&c, w, h);
#pragma omp for schedule(dynamic, 1)
        for (int y = 0; y < h; ++y) {
            for (int x = 0; x < w; ++x) {
                size_t p = (size_t)y * w + x;
                oray view;
                memcpy(view.o, c.cam_o, sizeof(float) * c.D);
                primary_dir(&c, (float)x, (float)y, view.d);
                ids_out[p] = -1;
                if (dist_out) dist_out[p] = 0;
                if (s->kind == NTR_SCENE_BOX) {
                    oray normal;
                    memset(&normal, 0, sizeof normal);
                    float d = hypercube_intersects(c.D, &view, &normal, FLT_MAX);
                    if (d) { ids_out[p] = 0; if (dist_out) dist_out[p] = d; }
                    continue;
                }
                float dist = aabb_distance(&c, &view);
                if (dist < 0) continue;
                otarget none = { NONE_REF, 0 };
                ohit hit;
                memset(&hit, 0, sizeof hit);
                hit.dist = FLT_MAX;
                ohits th = { NULL, 0, 0 };
                if (tree_intersects(&c, s->root, &view, none, &hit, &th, dist, FLT_MAX)) {
                    ids_out[p] = flat_id(s, hit.target);
                    if (dist_out) dist_out[p] = hit.dist;
                }
                free(th.v);
            }
        }
    }
    return 0;
}

/* screen_coord_to_ray, ntracer_body.hpp:3342-3358 */
ORACLE_API int oracle_screen_coord_to_ray(int dim, const float *cam_axes, float x, float y, int w, int h, float fov,
                                          float *dir_out) {
    ntr_scene_desc s;
    memset(&s, 0, sizeof s);
    s.dim = dim; s.fov = fov;
    octx c;
    ctx_init(&c, &s, NULL, cam_axes);
    set_view(&c, w, h);
    primary_dir(&c, x, y, dir_out);
    return 0;
}

/* Debug/parity hook: KDNode.intersects for ONE ray returning the full hit records the way the Python binding
 * lists them (transparent hits in list order, then the opaque hit): per hit dist, flat id, origin[D], normal[D]. */
ORACLE_API int oracle_trace_ray_full(const ntr_scene_desc *s, const float *origin, const float *dir, float t_near,
                                     float t_far, uint32_t skip_ref, int32_t skip_lane, int max_hits, float *dist_out,
                                     int32_t *id_out, float *origin_out, float *normal_out) {
    octx c;
    ctx_init(&c, s, NULL, NULL);
    oray ray;
    memcpy(ray.o, origin, sizeof(float) * c.D);
    memcpy(ray.d, dir, sizeof(float) * c.D);
    otarget skip = { skip_ref, skip_lane };
    ohit hit;
    memset(&hit, 0, sizeof hit);
    hit.dist = FLT_MAX;
    ohits th = { NULL, 0, 0 };
    int did = tree_intersects(&c, s->root, &ray, skip, &hit, &th, t_near, t_far);
    int n = 0;
    for (size_t i = 0; i < th.n && n < max_hits; ++i, ++n) {
        dist_out[n] = th.v[i].dist;
        id_out[n] = flat_id(s, th.v[i].target);
        memcpy(origin_out + (size_t)n * c.D, th.v[i].normal.o, sizeof(float) * c.D);
        memcpy(normal_out + (size_t)n * c.D, th.v[i].normal.d, sizeof(float) * c.D);
    }
    if (did && n < max_hits) {
        dist_out[n] = hit.dist;
        id_out[n] = flat_id(s, hit.target);
        memcpy(origin_out + (size_t)n * c.D, hit.normal.o, sizeof(float) * c.D);
        memcpy(normal_out + (size_t)n * c.D, hit.normal.d, sizeof(float) * c.D);
        ++n;
    }
    free(th.v);
    return n;
}
