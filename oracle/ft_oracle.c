/*
 * ft_oracle.c -- CPU restatement of FrayTracer's per-ray hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (fraytracer_b200/,
 * include/) may include, link or call this file; only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py use it, as the checker and as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (JanWosnitza/FrayTracer, F#) ships no tests,
 * golden images or fixtures (FrayTracer.sln:6-9 lists two projects, no test
 * project) and no .NET runtime exists in this environment, so this file could
 * not be checked against the reference's own outputs.  It is pinned only by
 * (a) known-answer tests derived by hand from the reference formulas and
 * (b) the published behaviour of System.Random / System.Numerics (.NET 6),
 * see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/FrayTracer).  Arithmetic is IEEE binary32, one rounding
 * per operation, no FMA contraction (compile with -ffp-contract=off), in the
 * operation order of the F# source / System.Numerics:
 *   Vector3.Dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z   (x64 dpps reduction)
 *   Length             = sqrt(Dot(v,v));  Distance(a,b) = Length(a-b)
 *   Normalize(v)       = v / Length(v)    (per-component IEEE divide)
 *   v / s, v * s       = per-component
 *   Vector3.Lerp(a,b,t)= a*(1-t) + b*t    (.NET 5+)
 *   MathF.Min/Max      = .NET Core 3.0+ semantics (NaN propagating)
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fraytracer_b200.h" /* scene description structs only */

#define FTO_API __attribute__((visibility("default")))

/* FTO_ASSUME: build-time switches for the three assumptions about code that lives OUTSIDE /root/reference
 * and that nothing reference-held pins.  0 (the default, what every parity claim in this repository means) = the .NET 6
 * in-box System.Numerics and Array.Sort.  tests/test_assumption_sensitivity.py builds one library per switch and reports how
 * many pixels of the reference scenes each assumption moves, so that a maintainer with .NET knows what to check first.
 *   bit 0  Vector3 / float32 and Vector3.Normalize multiply by the reciprocal (System.Numerics.Vectors 4.5.0 / .NET Core 2.x:
 *          `float invDiv = 1.0f / value2`) instead of dividing per component (.NET Core 3.0+)
 *   bit 1  Vector3.Dot reduces as x*x' + (y*y' + z*z') instead of (x*x' + y*y') + z*z' (dpps)
 *   bit 2  Array.sortInPlaceBy keeps equal LowerBounds in their original order (a stable sort) instead of the introsort's  */
#ifndef FTO_ASSUME
#define FTO_ASSUME 0
#endif
FTO_API int fto_assumptions(void) { return FTO_ASSUME; }

/* ------------------------------------------------------------------------ */
/* System.Numerics restatement                                               */
/* ------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdivv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
#if FTO_ASSUME & 1
static inline v3 vdiv(v3 a, float s) { float inv = 1.0f / s; return V(a.x * inv, a.y * inv, a.z * inv); }
#else
static inline v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
#endif
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
#if FTO_ASSUME & 2
static inline float vdot(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
#else
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
#endif
static inline float vlen2(v3 a) { return vdot(a, a); }
static inline float vlen(v3 a) { return sqrtf(vdot(a, a)); }
static inline float vdist(v3 a, v3 b) { return vlen(vsub(a, b)); }
static inline float vdist2(v3 a, v3 b) { return vlen2(vsub(a, b)); }
static inline v3 vnormalize(v3 a) { return vdiv(a, vlen(a)); }
static inline v3 vcross(v3 a, v3 b) {
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 vmin(v3 a, v3 b) {
    return V(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z);
}
static inline v3 vmax(v3 a, v3 b) {
    return V(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z);
}
static inline v3 vlerp(v3 a, v3 b, float t) { return vadd(vscale(a, 1.0f - t), vscale(b, t)); }
/* Math.fs:69  Vector3.inverseLength v = v / v.LengthSquared() */
static inline v3 vinverse_length(v3 a) { return vdiv(a, vlen2(a)); }

/* System.MathF.Max / Min (.NET Core 3.0+) */
static inline float net_max(float a, float b) {
    if (a != b) {
        if (!isnan(a)) return b < a ? a : b;
        return a;
    }
    return signbit(b) ? a : b;
}
static inline float net_min(float a, float b) {
    if (a != b) {
        if (!isnan(a)) return a < b ? a : b;
        return a;
    }
    return signbit(a) ? a : b;
}
/* Math.fs:51  clamp01 x = MathF.Max(0f, MathF.Min(1f, x)) */
static inline float clamp01(float x) { return net_max(0.0f, net_min(1.0f, x)); }
/* Math.fs:40  sign_i (NaN would throw in .NET; not reachable from finite input) */
static inline int sign_i(float x) { return (x > 0.0f) - (x < 0.0f); }
/* Math.fs:57-59  MathF.Floor/Ceiling |> int : x64 cvttss2si, out of range -> INT_MIN */
static inline int conv_i4(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int)x;
}
static inline int floor_i(float x) { return conv_i4(floorf(x)); }
static inline int ceiling_i(float x) { return conv_i4(ceilf(x)); }
static inline int clamp_i(int lo, int hi, int x) {
    int m = hi < x ? hi : x; /* Math.Min(max, x) */
    return lo > m ? lo : m;  /* Math.Max(min, .) */
}

/* ------------------------------------------------------------------------ */
/* System.Random(seed) -- Knuth subtractive generator, .NET seeded algorithm */
/* (needed to regenerate Program.fs's scene; Program.fs:14, Random.fs:9-40)  */
/* ------------------------------------------------------------------------ */
typedef struct fto_random {
    int32_t sa[56];
    int inext, inextp;
} fto_random;

FTO_API fto_random *fto_random_create(int32_t seed) {
    fto_random *r = (fto_random *)calloc(1, sizeof(fto_random));
    const int32_t MBIG = 2147483647, MSEED = 161803398;
    int32_t sub = (seed == INT32_MIN) ? INT32_MAX : (seed < 0 ? -seed : seed);
    int32_t mj = MSEED - sub;
    r->sa[55] = mj;
    int32_t mk = 1;
    int ii = 0;
    for (int i = 1; i < 55; i++) {
        ii += 21;
        if (ii >= 55) ii -= 55;
        r->sa[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += MBIG;
        mj = r->sa[ii];
    }
    for (int k = 1; k < 5; k++) {
        for (int i = 1; i < 56; i++) {
            int n = i + 30;
            if (n >= 55) n -= 55;
            r->sa[i] = (int32_t)((uint32_t)r->sa[i] - (uint32_t)r->sa[1 + n]);
            if (r->sa[i] < 0) r->sa[i] += MBIG;
        }
    }
    r->inext = 0;
    r->inextp = 21;
    return r;
}
FTO_API void fto_random_destroy(fto_random *r) { free(r); }
FTO_API int32_t fto_random_next(fto_random *r) { /* InternalSample */
    const int32_t MBIG = 2147483647;
    int a = r->inext, b = r->inextp;
    if (++a >= 56) a = 1;
    if (++b >= 56) b = 1;
    int32_t v = (int32_t)((uint32_t)r->sa[a] - (uint32_t)r->sa[b]);
    if (v == MBIG) v--;
    if (v < 0) v += MBIG;
    r->sa[a] = v;
    r->inext = a;
    r->inextp = b;
    return v;
}
FTO_API double fto_random_next_double(fto_random *r) {
    return fto_random_next(r) * (1.0 / 2147483647.0);
}

/* ------------------------------------------------------------------------ */
/* .NET 6 Array.Sort(keys, items) for float keys: NaNs to front + introsort. */
/* F# Array.sortInPlaceBy (SdfBoundary.fs:267-268) lands here: it is NOT     */
/* stable, so the order of equal LowerBounds is this algorithm's.            */
/* ------------------------------------------------------------------------ */
typedef struct { float lb; int32_t item; } gitem;

static inline void gi_swap(gitem *a, int i, int j) { gitem t = a[i]; a[i] = a[j]; a[j] = t; }
static inline void gi_swap_if_greater(gitem *a, int i, int j) {
    if (a[i].lb > a[j].lb) gi_swap(a, i, j);
}
static void gi_insertion_sort(gitem *a, int n) {
    for (int i = 0; i < n - 1; i++) {
        gitem t = a[i + 1];
        int j = i;
        while (j >= 0 && t.lb < a[j].lb) { a[j + 1] = a[j]; j--; }
        a[j + 1] = t;
    }
}
static void gi_down_heap(gitem *a, int i, int n) {
    gitem d = a[i - 1];
    while (i <= n >> 1) {
        int child = 2 * i;
        if (child < n && a[child - 1].lb < a[child].lb) child++;
        if (!(d.lb < a[child - 1].lb)) break;
        a[i - 1] = a[child - 1];
        i = child;
    }
    a[i - 1] = d;
}
static void gi_heap_sort(gitem *a, int n) {
    for (int i = n >> 1; i >= 1; i--) gi_down_heap(a, i, n);
    for (int i = n; i > 1; i--) { gi_swap(a, 0, i - 1); gi_down_heap(a, 1, i - 1); }
}
static int gi_pick_pivot_and_partition(gitem *a, int n) {
    int hi = n - 1, middle = hi >> 1;
    gi_swap_if_greater(a, 0, middle);
    gi_swap_if_greater(a, 0, hi);
    gi_swap_if_greater(a, middle, hi);
    float pivot = a[middle].lb;
    gi_swap(a, middle, hi - 1);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (pivot > a[++left].lb) {}
        while (pivot < a[--right].lb) {}
        if (left >= right) break;
        gi_swap(a, left, right);
    }
    if (left != hi - 1) gi_swap(a, left, hi - 1);
    return left;
}
static void gi_intro_sort(gitem *a, int n, int depth_limit) {
    int size = n;
    while (size > 1) {
        if (size <= 16) {
            if (size == 2) { gi_swap_if_greater(a, 0, 1); return; }
            if (size == 3) {
                gi_swap_if_greater(a, 0, 1);
                gi_swap_if_greater(a, 0, 2);
                gi_swap_if_greater(a, 1, 2);
                return;
            }
            gi_insertion_sort(a, size);
            return;
        }
        if (depth_limit == 0) { gi_heap_sort(a, size); return; }
        depth_limit--;
        int p = gi_pick_pivot_and_partition(a, size);
        gi_intro_sort(a + p + 1, size - (p + 1), depth_limit);
        size = p;
    }
}
static void dotnet_sort_by_lower_bound(gitem *a, int n) {
    if (n < 2) return;
#if FTO_ASSUME & 4
    gi_insertion_sort(a, n); /* stable: equal keys keep their order (strict < moves) */
    return;
#endif
    int left = 0;
    for (int i = 0; i < n; i++)
        if (isnan(a[i].lb)) { gi_swap(a, left, i); left++; }
    if (left == n) return;
    a += left; n -= left;
    int log2n = 0;
    for (unsigned v = (unsigned)n; v >>= 1;) log2n++;
    gi_intro_sort(a, n, 2 * (log2n + 1));
}
FTO_API void fto_dotnet_sort(float *keys, int32_t *items, int32_t n) { /* for tests */
    gitem *a = (gitem *)malloc(sizeof(gitem) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { a[i].lb = keys[i]; a[i].item = items[i]; }
    dotnet_sort_by_lower_bound(a, n);
    for (int i = 0; i < n; i++) { keys[i] = a[i].lb; items[i] = a[i].item; }
    free(a);
}

/* ------------------------------------------------------------------------ */
/* Boundaries  (Types.fs:19-24, SdfBoundary.fs:7-63)                         */
/* ------------------------------------------------------------------------ */
typedef struct { v3 c; float r; } bnd;

/* SdfBoundary.fs:7-22 */
static bnd bnd_union(bnd a, bnd b) {
    v3 diff = vsub(b.c, a.c);
    float distance = vlen(diff);
    if (distance + b.r <= a.r) return a;
    if (distance + a.r <= b.r) return b;
    v3 dir = vdiv(diff, distance);
    v3 a2 = vsub(a.c, vscale(dir, a.r));
    v3 b2 = vadd(b.c, vscale(dir, b.r));
    bnd r;
    r.c = vscale(vadd(a2, b2), 0.5f);
    r.r = vdist(a2, b2) * 0.5f;
    return r;
}
/* SdfBoundary.fs:29-49 (the radius formula is the reference's, kept verbatim) */
static bnd bnd_intersection(bnd a, bnd b) {
    v3 diff = vsub(b.c, a.c);
    float distance = vlen(diff);
    if (distance + b.r <= a.r) return b;
    if (distance + a.r <= b.r) return a;
    v3 dir = vdiv(diff, distance);
    v3 a2 = vadd(a.c, vscale(dir, a.r));
    v3 b2 = vsub(b.c, vscale(dir, b.r));
    bnd r;
    r.c = vscale(vadd(a2, b2), 0.5f);
    float d2 = distance * distance;
    float aR2 = a.r * a.r;
    float bR2 = b.r * b.r;
    r.r = sqrtf(4.0f * d2 * aR2 - (d2 - bR2 + aR2)) / (2.0f * distance);
    return r;
}
/* SdfBoundary.fs:62-63 */
static inline float bnd_min_distance(bnd b, v3 p) { return vdist(b.c, p) - b.r; }
static inline float bnd_max_distance(bnd b, v3 p) { return vdist(b.c, p) + b.r; }

/* ------------------------------------------------------------------------ */
/* Spatial lookup (SdfBoundary.fs:211-282)                                   */
/* ------------------------------------------------------------------------ */
typedef struct { v3 center; int32_t n; gitem *items; } gcell;
typedef struct grid {
    v3 aabb_min, cell_size, cell_size_inv;
    int cx, cy, cz;
    gcell *cells; /* cells[(x*cy + y)*cz + z] as the .NET [,,] array */
    int64_t n_entries;
    int64_t empty_cells;
} grid;

static grid *grid_build(const bnd *b, int n) {
    grid *g = (grid *)calloc(1, sizeof(grid));
    /* :229-230 */
    v3 aabb_min = vsub(b[0].c, V(b[0].r, b[0].r, b[0].r));
    v3 aabb_max = vadd(b[0].c, V(b[0].r, b[0].r, b[0].r));
    float sum = b[0].r;
    for (int i = 1; i < n; i++) {
        aabb_min = vmin(aabb_min, vsub(b[i].c, V(b[i].r, b[i].r, b[i].r)));
        aabb_max = vmax(aabb_max, vadd(b[i].c, V(b[i].r, b[i].r, b[i].r)));
        sum = sum + b[i].r; /* :232 Seq.reduce (+) */
    }
    float count_size = 1.5f * (sum / (float)n); /* :233 */
    v3 aabb_size = vsub(aabb_max, aabb_min);    /* :235 */
    /* :237-239  all three counts use aabbSize.X */
    int cx = ceiling_i(aabb_size.x / count_size); if (cx < 1) cx = 1;
    int cy = cx, cz = cx;
    v3 cell_size = vdivv(aabb_size, V((float)cx, (float)cy, (float)cz)); /* :240 */
    v3 cell_size_inv = vdivv(V(1.0f, 1.0f, 1.0f), cell_size);            /* :241 */
    g->aabb_min = aabb_min; g->cell_size = cell_size; g->cell_size_inv = cell_size_inv;
    g->cx = cx; g->cy = cy; g->cz = cz;
    g->cells = (gcell *)calloc((size_t)cx * cy * cz, sizeof(gcell));
    v3 half = vscale(cell_size, 0.5f);
    float half_len = vlen(half);
    gitem *tmp = (gitem *)malloc(sizeof(gitem) * (size_t)n);
    for (int x = 0; x < cx; x++)
        for (int y = 0; y < cy; y++)
            for (int z = 0; z < cz; z++) {
                /* :246 */
                v3 center = vadd(vadd(aabb_min, half), vmul(cell_size, V((float)x, (float)y, (float)z)));
                /* :248-253 */
                float m = bnd_max_distance(b[0], center);
                for (int i = 1; i < n; i++) {
                    float d = bnd_max_distance(b[i], center);
                    if (d < m) m = d; /* Seq.min */
                }
                float upper_bound = m + half_len;
                /* :255-264 */
                int k = 0;
                for (int i = 0; i < n; i++) {
                    float md = bnd_min_distance(b[i], center);
                    if (md < upper_bound) { tmp[k].lb = md; tmp[k].item = i; k++; }
                }
                dotnet_sort_by_lower_bound(tmp, k); /* :267-268 */
                gcell *c = &g->cells[((size_t)x * cy + y) * cz + z];
                if (k == 0) g->empty_cells++; /* the reference would index Items.[0] of an empty array */
                c->center = center;
                c->n = k;
                c->items = (gitem *)malloc(sizeof(gitem) * (size_t)(k > 0 ? k : 1));
                memcpy(c->items, tmp, sizeof(gitem) * (size_t)k);
                g->n_entries += k;
            }
    free(tmp);
    return g;
}
static void grid_free(grid *g) {
    if (!g) return;
    size_t nc = (size_t)g->cx * g->cy * g->cz;
    for (size_t i = 0; i < nc; i++) free(g->cells[i].items);
    free(g->cells);
    free(g);
}
/* :276-282 */
static __thread int tl_outside = 0;
static inline const gcell *grid_lookup(const grid *g, v3 p) {
    v3 c = vmul(vsub(p, g->aabb_min), g->cell_size_inv);
    int fx = floor_i(c.x), fy = floor_i(c.y), fz = floor_i(c.z);
    int ix = clamp_i(0, g->cx - 1, fx);
    int iy = clamp_i(0, g->cy - 1, fy);
    int iz = clamp_i(0, g->cz - 1, fz);
    tl_outside = (ix != fx) || (iy != fy) || (iz != fz);
    return &g->cells[((size_t)ix * g->cy + iy) * g->cz + iz];
}

/* ------------------------------------------------------------------------ */
/* Forms (SdfForm.fs)                                                        */
/* ------------------------------------------------------------------------ */
typedef struct onode {
    int kind;
    bnd boundary;
    /* primitives: precomputed closure captures */
    v3 a, b, c;            /* sphere: a=C | capsule: a=From,b=dir,c=dirInv | torus: a=C,b=N | tri: V1,V2,V3 */
    v3 e21, e32, e13;      /* triangle edges */
    v3 i21, i32, i13;      /* triangle inverseLength edges */
    v3 nor, n21, n32, n13; /* triangle normals */
    float r0, r1, r2;      /* sphere: r0=R | capsule: r0=R | torus: r0=planeD r1=Major r2=Minor | tri: r0=R | smooth: r0=strength r1=-1/strength */
    int nchild;
    struct onode **child;
    grid *g;
} onode;

/* counters of the reference algorithm's work (per thread, merged by caller) */
typedef struct fto_counters {
    int64_t march_steps;    /* Distance evaluations inside tryTrace loops    */
    int64_t primary_rays, primary_hits, shadow_rays;
    int64_t eval_sphere, eval_capsule, eval_torus, eval_triangle_edge, eval_triangle_face;
    int64_t union_lookups;  /* grid cell lookups (form unions + material unions) */
    int64_t lower_bound_tests, bound_tests; /* the two pruning tests          */
    int64_t intersect_tests;
    int64_t max_steps;      /* longest single march                           */
    int64_t outside_lookups, outside_bound_tests, outside_evals; /* same, for positions outside the grid AABB */
} fto_counters;

static __thread fto_counters *tl_ctr = NULL;
#define CNT(field) do { if (tl_ctr) tl_ctr->field++; } while (0)

static float form_distance(const onode *n, v3 p);

/* SdfForm.fs:22-34 */
static float union_distance(const onode *n, v3 p) {
    const gcell *cell = grid_lookup(n->g, p);
    const int outside = tl_outside;
    CNT(union_lookups);
    if (outside) CNT(outside_lookups);
    float dtc = vdist(p, cell->center);
    float min = form_distance(n->child[cell->items[0].item], p);
    for (int i = 1; i < cell->n; i++) {
        const gitem *it = &cell->items[i];
        const onode *f = n->child[it->item];
        CNT(lower_bound_tests);
        if (min > it->lb - dtc) {
            CNT(bound_tests);
            if (outside) CNT(outside_bound_tests);
            if (min > bnd_min_distance(f->boundary, p)) {
                if (outside) CNT(outside_evals);
                min = net_min(min, form_distance(f, p));
            }
        }
    }
    return min;
}

static float form_distance(const onode *n, v3 p) {
    switch (n->kind) {
    case FT_FORM_SPHERE: /* SdfForm.fs:127-129 */
        CNT(eval_sphere);
        return vdist(n->a, p) - n->r0;
    case FT_FORM_CAPSULE: { /* SdfForm.fs:151-164 */
        CNT(eval_capsule);
        v3 diff = vsub(p, n->a);
        float t = vdot(diff, n->c);
        float d;
        if (t <= 0.0f) d = vlen(diff);
        else if (t >= 1.0f) d = vdist(diff, n->b);
        else d = vdist(diff, vscale(n->b, t));
        return d - n->r0;
    }
    case FT_FORM_TORUS: { /* SdfForm.fs:189-194 */
        CNT(eval_torus);
        float dplane = vdot(p, n->b) + n->r0;
        float dcenter = vdist(n->a, vsub(p, vscale(n->b, dplane)));
        float dcircle = dcenter - n->r1;
        return sqrtf(dplane * dplane + dcircle * dcircle) - n->r2; /* Vector2.Length */
    }
    case FT_FORM_CONE: { /* EXTENSION, no reference implementation: exact capped-cone distance.
                            a = From, b = To - From, r0 = ra, r1 = rb, r2 = baba */
        CNT(eval_capsule);
        v3 pa = vsub(p, n->a);
        float ra = n->r0, rb = n->r1, baba = n->r2, rba = rb - ra, k = rba * rba + baba;
        float papa = vdot(pa, pa);
        float paba = vdot(pa, n->b) / baba;
        float x2 = papa - (paba * paba) * baba;
        float x = sqrtf(x2 > 0.0f ? x2 : 0.0f);
        float cx = x - (paba < 0.5f ? ra : rb);
        float cax = cx > 0.0f ? cx : 0.0f;
        float cay = fabsf(paba - 0.5f) - 0.5f;
        float fr = (rba * (x - ra) + paba * baba) / k;
        float f = fr < 0.0f ? 0.0f : (fr > 1.0f ? 1.0f : fr);
        float cbx = (x - ra) - f * rba;
        float cby = paba - f;
        float da = cax * cax + (cay * cay) * baba, db = cbx * cbx + (cby * cby) * baba;
        float dm = sqrtf(da < db ? da : db);
        return (cbx < 0.0f && cay < 0.0f) ? -dm : dm;
    }
    case FT_FORM_TRIANGLE: { /* SdfForm.fs:227-250 */
        v3 p1 = vsub(p, n->a), p2 = vsub(p, n->b), p3 = vsub(p, n->c);
        float d;
        if (sign_i(vdot(n->n21, p1)) + sign_i(vdot(n->n32, p2)) + sign_i(vdot(n->n13, p3)) < 2) {
            CNT(eval_triangle_edge);
            float d21 = vdist2(p1, vscale(n->e21, clamp01(vdot(n->i21, p1))));
            float d32 = vdist2(p2, vscale(n->e32, clamp01(vdot(n->i32, p2))));
            float d13 = vdist2(p3, vscale(n->e13, clamp01(vdot(n->i13, p3))));
            d = sqrtf(net_min(d13, net_min(d32, d21)));
        } else {
            CNT(eval_triangle_face);
            d = fabsf(vdot(n->nor, p1));
        }
        return d - n->r0;
    }
    case FT_FORM_UNION:
        return union_distance(n, p);
    case FT_FORM_SUBTRACT: { /* SdfForm.fs:45-47: MathF.Max(-(b p), a p) */
        float a = form_distance(n->child[0], p);
        float b = form_distance(n->child[1], p);
        return net_max(-b, a);
    }
    case FT_FORM_INTERSECT: { /* SdfForm.fs:58-64 */
        float max = form_distance(n->child[0], p);
        for (int i = 1; i < n->nchild; i++) {
            CNT(intersect_tests);
            if (max < bnd_max_distance(n->child[i]->boundary, p))
                max = net_max(max, form_distance(n->child[i], p));
        }
        return max;
    }
    case FT_FORM_UNION_SMOOTH: { /* SdfForm.fs:75-82 */
        float sum = 0.0f;
        for (int i = 0; i < n->nchild; i++) sum = sum + expf(n->r1 * form_distance(n->child[i], p));
        return -logf(sum) * n->r0;
    }
    }
    return NAN;
}

/* ------------------------------------------------------------------------ */
/* Objects (SdfObject.fs)                                                    */
/* ------------------------------------------------------------------------ */
typedef struct oobj {
    int kind;
    onode *form;      /* the object's Form                                    */
    int material;     /* LEAF                                                 */
    int leaf_id;      /* index in the description's object table              */
    struct oobj *base; /* SUBTRACT / INTERSECT: the object whose material is kept */
    int nchild;
    struct oobj **child; /* UNION */
    grid *g;          /* UNION: the second, independently built lookup (SdfObject.fs:26) */
} oobj;

/* SdfObject.fs:27-46: returns the leaf object whose material colours p */
static const oobj *object_material_leaf(const oobj *o, v3 p) {
    for (;;) {
        if (o->kind == FT_OBJECT_LEAF) return o;
        if (o->kind == FT_OBJECT_SUBTRACT || o->kind == FT_OBJECT_INTERSECT) { o = o->base; continue; }
        const gcell *cell = grid_lookup(o->g, p);
        CNT(union_lookups);
        const oobj *material = o->child[cell->items[0].item];
        float min = form_distance(material->form, p);
        float dtc = vdist(p, cell->center);
        for (int i = 0; i < cell->n; i++) {
            const gitem *it = &cell->items[i];
            const oobj *c = o->child[it->item];
            CNT(lower_bound_tests);
            if (min > it->lb - dtc) {
                CNT(bound_tests);
                if (min > bnd_min_distance(c->form->boundary, p)) {
                    float d = form_distance(c->form, p);
                    if (d < min) { min = d; material = c; }
                }
            }
        }
        o = material; /* material |> SdfMaterial.getColor: recurse into the child's Material */
    }
}

/* ------------------------------------------------------------------------ */
/* Scene                                                                     */
/* ------------------------------------------------------------------------ */
typedef struct olight {
    int kind;
    v3 v;      /* directional: normalize(-direction) (SdfLight.fs:7); point: position */
    v3 color;
} olight;

typedef struct fto_scene {
    int n_forms, n_objects, n_lights, n_materials;
    onode **forms;   /* resolved node of each described form (single-child composites alias the child) */
    onode **owned;   /* every allocated node */
    int n_owned, cap_owned;
    oobj **objects;
    oobj **owned_obj;
    int n_owned_obj;
    olight *lights;
    v3 *material_rgb;
    int *material_kind;
    v3 background;
    oobj *root;
    grid **grids; int n_grids, cap_grids; /* form-union grids in creation order, for tests */
    int64_t step_cap;
} fto_scene;

typedef struct { v3 origin, direction; float length, epsilon; } oray;

/* Ray.fs:6-13 */
static inline v3 ray_get(const oray *r, float l) { return vadd(r->origin, vscale(r->direction, l)); }
static inline oray ray_move(oray r, float l) {
    r.origin = ray_get(&r, l);
    r.length = r.length - l;
    return r;
}

/* SdfForm.fs:93-104.  Returns 1 on hit; *hit is the (un-moved) hit ray.      */
static int form_try_trace(const fto_scene *s, const onode *f, oray ray, oray *hit, int64_t *steps_out) {
    int64_t steps = 0;
    int result = 0;
    for (;;) {
        if (ray.length <= 0.0f) break;
        if (steps >= s->step_cap) break; /* watchdog, not in the reference */
        float d = form_distance(f, ray.origin);
        steps++;
        if (d < ray.epsilon) { *hit = ray; result = 1; break; }
        ray = ray_move(ray, d);
    }
    if (tl_ctr) {
        tl_ctr->march_steps += steps;
        if (steps > tl_ctr->max_steps) tl_ctr->max_steps = steps;
    }
    if (steps_out) *steps_out = steps;
    return result;
}

/* SdfForm.fs:106-115 */
static v3 form_normal_from_ray(const onode *f, const oray *ray) {
    float e = ray->epsilon * 0.125f;
    v3 p = ray_get(ray, -ray->epsilon);
    float dx = form_distance(f, V(p.x + e, p.y, p.z));
    float dy = form_distance(f, V(p.x, p.y + e, p.z));
    float dz = form_distance(f, V(p.x, p.y, p.z + e));
    float d0 = form_distance(f, p);
    return vnormalize(vsub(V(dx, dy, dz), V(d0, d0, d0)));
}

typedef struct { oray ray; v3 normal; v3 color; int leaf; } otrace;

/* SdfObject.fs:66-78.  `shade` = 0 skips normal+colour (pure, discarded by
 * the light closures, SdfLight.fs:17-20,38-41).                             */
static int object_try_trace(const fto_scene *s, const oobj *o, oray ray, int shade, otrace *out, int64_t *steps) {
    oray hit;
    if (!form_try_trace(s, o->form, ray, &hit, steps)) return 0;
    if (shade) {
        out->normal = form_normal_from_ray(o->form, &hit);
        out->ray = ray_move(hit, -ray.epsilon);
        const oobj *leaf = object_material_leaf(o, hit.origin);
        out->leaf = leaf->leaf_id;
        out->color = s->material_rgb[leaf->material];
    }
    return 1;
}

typedef struct { v3 color; int leaf; int64_t steps; uint32_t shadow; } opixel;

/* SdfScene.fs:7-28 */
static opixel scene_trace(const fto_scene *s, oray ray) {
    opixel px;
    otrace res;
    px.shadow = 0;
    CNT(primary_rays);
    if (!object_try_trace(s, s->root, ray, 1, &res, &px.steps)) {
        px.color = s->background;
        px.leaf = -1;
        return px;
    }
    CNT(primary_hits);
    v3 light_color = s->background;
    for (int l = 0; l < s->n_lights; l++) {
        const olight *lt = &s->lights[l];
        v3 position = res.ray.origin; /* result.Position (Types.fs:65) */
        v3 ldir = (lt->kind == FT_LIGHT_DIRECTIONAL) ? lt->v : vnormalize(vsub(lt->v, position));
        float lcos = vdot(res.normal, ldir);
        if (lcos > 0.0f) {
            oray sr;
            v3 intensity;
            sr.origin = res.ray.origin;
            sr.epsilon = res.ray.epsilon;
            if (lt->kind == FT_LIGHT_DIRECTIONAL) { /* SdfLight.fs:10-20 */
                sr.direction = lt->v;
                sr.length = 1000.0f;
                intensity = lt->color;
            } else { /* SdfLight.fs:26-41 */
                v3 diff = vsub(lt->v, res.ray.origin);
                float d2 = vlen2(diff);
                float d = sqrtf(d2);
                sr.direction = vdiv(diff, d2);
                sr.length = d;
                intensity = vdiv(lt->color, d2);
            }
            CNT(shadow_rays);
            otrace dummy;
            if (!object_try_trace(s, s->root, sr, 0, &dummy, NULL)) {
                light_color = vadd(light_color, vscale(intensity, lcos));
                px.shadow |= (1u << l);
            }
        }
    }
    const float pi_inv = 1.0f / 3.14159274f; /* Math.fs:30  1f / MathF.PI */
    px.color = vmul(res.color, vscale(light_color, pi_inv));
    px.leaf = res.leaf;
    return px;
}

/* ------------------------------------------------------------------------ */
/* Scene construction from the flat description                              */
/* ------------------------------------------------------------------------ */
static void set_err(char *err, int errlen, const char *msg) {
    if (err && errlen > 0) { strncpy(err, msg, (size_t)errlen - 1); err[errlen - 1] = 0; }
}
static onode *new_node(fto_scene *s) {
    if (s->n_owned == s->cap_owned) {
        s->cap_owned = s->cap_owned ? s->cap_owned * 2 : 64;
        s->owned = (onode **)realloc(s->owned, sizeof(onode *) * (size_t)s->cap_owned);
    }
    onode *n = (onode *)calloc(1, sizeof(onode));
    s->owned[s->n_owned++] = n;
    return n;
}
static void add_grid(fto_scene *s, grid *g) {
    if (s->n_grids == s->cap_grids) {
        s->cap_grids = s->cap_grids ? s->cap_grids * 2 : 8;
        s->grids = (grid **)realloc(s->grids, sizeof(grid *) * (size_t)s->cap_grids);
    }
    s->grids[s->n_grids++] = g;
}

/* SdfForm.union / intersect / unionSmooth / subtract over already built children */
static onode *make_composite(fto_scene *s, int kind, onode **kids, int n, float strength, char *err, int errlen) {
    if (kind != FT_FORM_SUBTRACT) {
        if (n == 0) { set_err(err, errlen, "No SdfObjects given."); return NULL; } /* SdfForm.fs:16,53,71 */
        if (n == 1) return kids[0];                                                /* SdfForm.fs:17,54,72 */
    } else if (n != 2) { set_err(err, errlen, "subtract needs exactly two forms"); return NULL; }
    onode *nd = new_node(s);
    nd->kind = kind;
    nd->nchild = n;
    nd->child = (onode **)malloc(sizeof(onode *) * (size_t)n);
    memcpy(nd->child, kids, sizeof(onode *) * (size_t)n);
    if (kind == FT_FORM_SUBTRACT) {
        nd->boundary = kids[0]->boundary; /* SdfForm.fs:48 */
    } else if (kind == FT_FORM_INTERSECT) {
        bnd b = kids[0]->boundary;        /* SdfForm.fs:66, Seq.reduce */
        for (int i = 1; i < n; i++) b = bnd_intersection(b, kids[i]->boundary);
        nd->boundary = b;
    } else {
        bnd b = kids[0]->boundary;        /* SdfForm.fs:36-39 / 87-90, Seq.reduce */
        for (int i = 1; i < n; i++) b = bnd_union(b, kids[i]->boundary);
        nd->boundary = b;
        if (kind == FT_FORM_UNION) {
            bnd *bb = (bnd *)malloc(sizeof(bnd) * (size_t)n);
            for (int i = 0; i < n; i++) bb[i] = kids[i]->boundary;
            nd->g = grid_build(bb, n); /* SdfForm.fs:19 */
            free(bb);
            add_grid(s, nd->g);
            if (nd->g->empty_cells) {
                set_err(err, errlen, "spatial lookup cell without candidates (non-finite boundary?): the reference "
                                     "would throw IndexOutOfRangeException in SdfForm.union");
                return NULL;
            }
        } else {
            nd->r0 = strength;
            nd->r1 = -1.0f / strength; /* SdfForm.fs:75 */
        }
    }
    return nd;
}

static onode *build_form(fto_scene *s, const ft_scene_desc *d, int id, int depth, char *err, int errlen) {
    if (id < 0 || id >= d->n_forms) { set_err(err, errlen, "form id out of range"); return NULL; }
    if (s->forms[id]) return s->forms[id];
    if (depth > 64) { set_err(err, errlen, "form nesting too deep (cycle?)"); return NULL; }
    const ft_form_desc *f = &d->forms[id];
    const float *p = f->p;
    onode *nd = NULL;
    switch (f->kind) {
    case FT_FORM_SPHERE: /* SdfForm.fs:125-135 */
        nd = new_node(s); nd->kind = f->kind;
        nd->a = V(p[0], p[1], p[2]); nd->r0 = p[3];
        nd->boundary.c = nd->a; nd->boundary.r = nd->r0;
        break;
    case FT_FORM_CAPSULE: { /* SdfForm.fs:145-170 */
        nd = new_node(s); nd->kind = f->kind;
        v3 from = V(p[0], p[1], p[2]), to = V(p[3], p[4], p[5]);
        nd->a = from; nd->b = vsub(to, from); nd->c = vinverse_length(nd->b); nd->r0 = p[6];
        nd->boundary.c = vlerp(from, to, 0.5f);
        nd->boundary.r = p[6] + vdist(from, to) * 0.5f;
        break;
    }
    case FT_FORM_TORUS: { /* SdfForm.fs:181-203 */
        nd = new_node(s); nd->kind = f->kind;
        nd->a = V(p[0], p[1], p[2]);
        nd->b = vnormalize(V(p[3], p[4], p[5]));
        nd->r0 = -(vdot(nd->a, nd->b));
        nd->r1 = p[6]; nd->r2 = p[7];
        nd->boundary.c = nd->a; nd->boundary.r = p[6] + p[7];
        break;
    }
    case FT_FORM_CONE: { /* EXTENSION (no reference) */
        nd = new_node(s); nd->kind = f->kind;
        v3 from = V(p[0], p[1], p[2]), to = V(p[3], p[4], p[5]);
        nd->a = from; nd->b = vsub(to, from); nd->r0 = p[6]; nd->r1 = p[7]; nd->r2 = vdot(nd->b, nd->b);
        if (!(nd->r2 > 0.0f) || !(p[6] >= 0.0f) || !(p[7] >= 0.0f)) { set_err(err, errlen, "cone: degenerate"); return NULL; }
        float h = vdist(from, to) * 0.5f, rm = p[6] > p[7] ? p[6] : p[7];
        nd->boundary.c = vlerp(from, to, 0.5f);
        nd->boundary.r = sqrtf(h * h + rm * rm) * 1.000001f;
        break;
    }
    case FT_FORM_TRIANGLE: { /* SdfForm.fs:214-263 */
        nd = new_node(s); nd->kind = f->kind;
        v3 v1 = V(p[0], p[1], p[2]), v2 = V(p[3], p[4], p[5]), v3_ = V(p[6], p[7], p[8]);
        float R = p[9];
        nd->a = v1; nd->b = v2; nd->c = v3_; nd->r0 = R;
        nd->e21 = vsub(v2, v1); nd->i21 = vinverse_length(nd->e21);
        nd->e32 = vsub(v3_, v2); nd->i32 = vinverse_length(nd->e32);
        nd->e13 = vsub(v1, v3_); nd->i13 = vinverse_length(nd->e13);
        nd->nor = vnormalize(vcross(nd->e21, nd->e13));
        nd->n21 = vnormalize(vcross(nd->e21, nd->nor));
        nd->n32 = vnormalize(vcross(nd->e32, nd->nor));
        nd->n13 = vnormalize(vcross(nd->e13, nd->nor));
        float area_inv = 0.5f / vlen2(vcross(vsub(v1, v2), vsub(v2, v3_)));
        float w1 = vlen2(vsub(v2, v3_)) * vdot(vsub(v1, v2), vsub(v1, v3_)) * area_inv;
        float w2 = vlen2(vsub(v1, v3_)) * vdot(vsub(v2, v1), vsub(v2, v3_)) * area_inv;
        float w3 = 1.0f - w1 - w2;
        nd->boundary.c = vadd(vadd(vscale(v1, w1), vscale(v2, w2)), vscale(v3_, w3));
        nd->boundary.r = vlen(nd->e21) * vlen(nd->e32) * vlen(nd->e13) / 2.0f / vlen(vcross(nd->e21, nd->e32)) + R;
        break;
    }
    case FT_FORM_UNION: case FT_FORM_SUBTRACT: case FT_FORM_INTERSECT: case FT_FORM_UNION_SMOOTH: {
        if (f->child_count < 0 || f->first_child < 0 || f->first_child + f->child_count > d->n_children) {
            set_err(err, errlen, "form child range out of bounds"); return NULL;
        }
        onode **kids = (onode **)malloc(sizeof(onode *) * (size_t)(f->child_count + 1));
        for (int i = 0; i < f->child_count; i++) {
            kids[i] = build_form(s, d, d->children[f->first_child + i], depth + 1, err, errlen);
            if (!kids[i]) { free(kids); return NULL; }
        }
        nd = make_composite(s, f->kind, kids, f->child_count, p[0], err, errlen);
        free(kids);
        if (!nd) return NULL;
        break;
    }
    default:
        set_err(err, errlen, "unsupported form kind");
        return NULL;
    }
    s->forms[id] = nd;
    return nd;
}

static oobj *build_object(fto_scene *s, const ft_scene_desc *d, int id, int depth, char *err, int errlen) {
    if (id < 0 || id >= d->n_objects) { set_err(err, errlen, "object id out of range"); return NULL; }
    if (s->objects[id]) return s->objects[id];
    if (depth > 64) { set_err(err, errlen, "object nesting too deep (cycle?)"); return NULL; }
    const ft_object_desc *od = &d->objects[id];
    if (od->kind != FT_OBJECT_LEAF &&
        (od->child_count < 0 || od->first_child < 0 || od->first_child + od->child_count > d->n_children)) {
        set_err(err, errlen, "object child range out of bounds"); return NULL;
    }
    oobj *o = (oobj *)calloc(1, sizeof(oobj));
    s->owned_obj[s->n_owned_obj++] = o;
    o->kind = od->kind;
    o->leaf_id = id;
    switch (od->kind) {
    case FT_OBJECT_LEAF: /* SdfObject.fs:6-10 */
        if (od->material < 0 || od->material >= d->n_materials) { set_err(err, errlen, "material id out of range"); return NULL; }
        o->form = build_form(s, d, od->form, 0, err, errlen);
        if (!o->form) return NULL;
        o->material = od->material;
        break;
    case FT_OBJECT_UNION: { /* SdfObject.fs:12-48 */
        if (od->child_count == 0) { set_err(err, errlen, "No SdfObjects given."); return NULL; }
        oobj **kids = (oobj **)malloc(sizeof(oobj *) * (size_t)od->child_count);
        for (int i = 0; i < od->child_count; i++) {
            kids[i] = build_object(s, d, d->children[od->first_child + i], depth + 1, err, errlen);
            if (!kids[i]) { free(kids); return NULL; }
        }
        if (od->child_count == 1) { /* SdfObject.fs:16: the object itself */
            oobj *only = kids[0];
            free(kids);
            s->objects[id] = only;
            return only;
        }
        o->nchild = od->child_count;
        o->child = kids;
        onode **fk = (onode **)malloc(sizeof(onode *) * (size_t)o->nchild);
        bnd *bb = (bnd *)malloc(sizeof(bnd) * (size_t)o->nchild);
        for (int i = 0; i < o->nchild; i++) { fk[i] = kids[i]->form; bb[i] = kids[i]->form->boundary; }
        o->form = make_composite(s, FT_FORM_UNION, fk, o->nchild, 0.0f, err, errlen); /* :18-21 */
        o->g = grid_build(bb, o->nchild);                                              /* :26    */
        free(fk); free(bb);
        if (!o->form) return NULL;
        break;
    }
    case FT_OBJECT_SUBTRACT: { /* SdfObject.fs:50-54 */
        if (od->child_count != 1) { set_err(err, errlen, "subtract object needs one base object"); return NULL; }
        o->base = build_object(s, d, d->children[od->first_child], depth + 1, err, errlen);
        if (!o->base) return NULL;
        onode *kids[2];
        kids[0] = o->base->form;
        kids[1] = build_form(s, d, od->form, 0, err, errlen);
        if (!kids[1]) return NULL;
        o->form = make_composite(s, FT_FORM_SUBTRACT, kids, 2, 0.0f, err, errlen);
        if (!o->form) return NULL;
        break;
    }
    case FT_OBJECT_INTERSECT: { /* SdfObject.fs:56-64 */
        if (od->child_count < 1) { set_err(err, errlen, "intersect object needs a base object"); return NULL; }
        o->base = build_object(s, d, d->children[od->first_child], depth + 1, err, errlen);
        if (!o->base) return NULL;
        onode **kids = (onode **)malloc(sizeof(onode *) * (size_t)od->child_count);
        kids[0] = o->base->form;
        for (int i = 1; i < od->child_count; i++) {
            kids[i] = build_form(s, d, d->children[od->first_child + i], 0, err, errlen);
            if (!kids[i]) { free(kids); return NULL; }
        }
        o->form = make_composite(s, FT_FORM_INTERSECT, kids, od->child_count, 0.0f, err, errlen);
        free(kids);
        if (!o->form) return NULL;
        break;
    }
    default:
        set_err(err, errlen, "unsupported object kind");
        return NULL;
    }
    s->objects[id] = o;
    return o;
}

FTO_API void fto_scene_destroy(fto_scene *s) {
    if (!s) return;
    for (int i = 0; i < s->n_owned; i++) {
        if (s->owned[i]->g) grid_free(s->owned[i]->g);
        free(s->owned[i]->child);
        free(s->owned[i]);
    }
    for (int i = 0; i < s->n_owned_obj; i++) {
        if (s->owned_obj[i]->g) grid_free(s->owned_obj[i]->g);
        free(s->owned_obj[i]->child);
        free(s->owned_obj[i]);
    }
    free(s->owned); free(s->owned_obj); free(s->forms); free(s->objects);
    free(s->lights); free(s->material_rgb); free(s->material_kind); free(s->grids);
    free(s);
}

FTO_API int fto_scene_create(const ft_scene_desc *d, fto_scene **out, char *err, int errlen) {
    if (!d || !out) { set_err(err, errlen, "null argument"); return FT_ERR_INVALID_ARGUMENT; }
    fto_scene *s = (fto_scene *)calloc(1, sizeof(fto_scene));
    s->n_forms = d->n_forms; s->n_objects = d->n_objects;
    s->n_lights = d->n_lights; s->n_materials = d->n_materials;
    s->forms = (onode **)calloc((size_t)d->n_forms + 1, sizeof(onode *));
    s->objects = (oobj **)calloc((size_t)d->n_objects + 1, sizeof(oobj *));
    s->owned_obj = (oobj **)calloc((size_t)d->n_objects + 1, sizeof(oobj *));
    s->material_rgb = (v3 *)calloc((size_t)d->n_materials + 1, sizeof(v3));
    s->material_kind = (int *)calloc((size_t)d->n_materials + 1, sizeof(int));
    s->lights = (olight *)calloc((size_t)d->n_lights + 1, sizeof(olight));
    s->step_cap = (int64_t)1 << 22;
    for (int i = 0; i < d->n_materials; i++) {
        s->material_rgb[i] = V(d->materials[i].rgb[0], d->materials[i].rgb[1], d->materials[i].rgb[2]);
        s->material_kind[i] = d->materials[i].kind;
    }
    for (int i = 0; i < d->n_lights; i++) {
        const ft_light_desc *l = &d->lights[i];
        s->lights[i].kind = l->kind;
        s->lights[i].color = V(l->rgb[0], l->rgb[1], l->rgb[2]);
        if (l->kind == FT_LIGHT_DIRECTIONAL) s->lights[i].v = vnormalize(vneg(V(l->v[0], l->v[1], l->v[2]))); /* SdfLight.fs:7 */
        else if (l->kind == FT_LIGHT_POINT) s->lights[i].v = V(l->v[0], l->v[1], l->v[2]);
        else { set_err(err, errlen, "unsupported light kind"); fto_scene_destroy(s); return FT_ERR_UNSUPPORTED_NODE; }
    }
    s->background = V(d->background[0], d->background[1], d->background[2]);
    char local[256] = {0};
    /* every constructor of the description runs, in index order (forms, then objects): an
     * empty composite fails at construction like the reference's failwith */
    int ok = 1;
    for (int i = 0; ok && i < d->n_forms; i++) ok = build_form(s, d, i, 0, local, (int)sizeof(local)) != NULL;
    for (int i = 0; ok && i < d->n_objects; i++) ok = build_object(s, d, i, 0, local, (int)sizeof(local)) != NULL;
    s->root = ok ? build_object(s, d, d->root_object, 0, local, (int)sizeof(local)) : NULL;
    if (!s->root) {
        set_err(err, errlen, local);
        int code = strstr(local, "No SdfObjects") ? FT_ERR_EMPTY_COMPOSITE
                 : strstr(local, "unsupported") ? FT_ERR_UNSUPPORTED_NODE : FT_ERR_INVALID_ARGUMENT;
        fto_scene_destroy(s);
        return code;
    }
    *out = s;
    return FT_OK;
}

FTO_API void fto_scene_set_step_cap(fto_scene *s, int64_t cap) { s->step_cap = cap; }

/* ---- grid introspection (form-union grids in creation order) ------------- */
FTO_API int fto_grid_count(const fto_scene *s) { return s->n_grids; }
FTO_API int fto_grid_info(const fto_scene *s, int gi, ft_grid_info *out) {
    if (gi < 0 || gi >= s->n_grids) return FT_ERR_INVALID_ARGUMENT;
    const grid *g = s->grids[gi];
    out->node = -1;
    out->count[0] = g->cx; out->count[1] = g->cy; out->count[2] = g->cz;
    out->aabb_min[0] = g->aabb_min.x; out->aabb_min[1] = g->aabb_min.y; out->aabb_min[2] = g->aabb_min.z;
    out->cell_size[0] = g->cell_size.x; out->cell_size[1] = g->cell_size.y; out->cell_size[2] = g->cell_size.z;
    out->cell_size_inv[0] = g->cell_size_inv.x; out->cell_size_inv[1] = g->cell_size_inv.y; out->cell_size_inv[2] = g->cell_size_inv.z;
    out->n_entries = g->n_entries;
    return FT_OK;
}
FTO_API int fto_grid_read(const fto_scene *s, int gi, float *centers, int64_t *begin, float *lb, int32_t *item) {
    if (gi < 0 || gi >= s->n_grids) return FT_ERR_INVALID_ARGUMENT;
    const grid *g = s->grids[gi];
    size_t nc = (size_t)g->cx * g->cy * g->cz;
    int64_t off = 0;
    for (size_t c = 0; c < nc; c++) {
        if (centers) { centers[3 * c] = g->cells[c].center.x; centers[3 * c + 1] = g->cells[c].center.y; centers[3 * c + 2] = g->cells[c].center.z; }
        if (begin) begin[c] = off;
        for (int i = 0; i < g->cells[c].n; i++) {
            if (lb) lb[off + i] = g->cells[c].items[i].lb;
            if (item) item[off + i] = g->cells[c].items[i].item;
        }
        off += g->cells[c].n;
    }
    if (begin) begin[nc] = off;
    return FT_OK;
}

/* ---- camera (Camera.fs) --------------------------------------------------- */
FTO_API float fto_lens_create(float fov) { return (float)sin((double)(fov * 0.5f)); } /* Camera.fs:13 */
FTO_API void fto_camera_look_at(const float pos[3], const float look_at[3], const float up[3], float nps, ft_camera *out) {
    v3 position = V(pos[0], pos[1], pos[2]);
    v3 forward = vnormalize(vsub(V(look_at[0], look_at[1], look_at[2]), position)); /* Camera.fs:34 */
    v3 right = vnormalize(vcross(V(up[0], up[1], up[2]), forward));                 /* Camera.fs:35 */
    v3 ups = vscale(vcross(forward, right), nps);                                   /* Camera.fs:40 */
    v3 rs = vscale(right, nps);                                                     /* Camera.fs:41 */
    out->position[0] = position.x; out->position[1] = position.y; out->position[2] = position.z;
    out->forward[0] = forward.x; out->forward[1] = forward.y; out->forward[2] = forward.z;
    out->up_scaled[0] = ups.x; out->up_scaled[1] = ups.y; out->up_scaled[2] = ups.z;
    out->right_scaled[0] = rs.x; out->right_scaled[1] = rs.y; out->right_scaled[2] = rs.z;
}

/* Image.fs:17-23 + Camera.fs:44-54 */
static oray pixel_ray(const ft_camera *cam, int x, int y, int W, int H, float eps, float len) {
    float max_size = (float)(W > H ? W : H);
    float u = (float)x / max_size, v = (float)y / max_size;
    v3 fwd = V(cam->forward[0], cam->forward[1], cam->forward[2]);
    v3 rs = V(cam->right_scaled[0], cam->right_scaled[1], cam->right_scaled[2]);
    v3 us = V(cam->up_scaled[0], cam->up_scaled[1], cam->up_scaled[2]);
    oray r;
    r.origin = V(cam->position[0], cam->position[1], cam->position[2]);
    r.direction = vnormalize(vadd(vadd(fwd, vscale(rs, u - 0.5f)), vscale(us, v - 0.5f)));
    r.epsilon = eps;
    r.length = len;
    return r;
}

/* ---- parallel drivers ------------------------------------------------------ */
typedef struct {
    const fto_scene *s;
    const ft_camera *cam;
    int W, H;
    float eps, len;
    int64_t n;
    const ft_ray *rays;
    const float *points;
    float *out_rgb; int32_t *out_leaf; int32_t *out_steps; uint32_t *out_shadow; float *out_dist;
    volatile int64_t next;
    int64_t start;
    int64_t chunk;
    int x0, xstride; /* render: column of item i is x0 + i*xstride */
    int mode; /* 0 render, 1 rays, 2 distance */
    int want_counters;
    pthread_mutex_t mu;
    fto_counters total;
} job;

static void merge_counters(fto_counters *t, const fto_counters *c) {
    t->march_steps += c->march_steps; t->primary_rays += c->primary_rays;
    t->primary_hits += c->primary_hits; t->shadow_rays += c->shadow_rays;
    t->eval_sphere += c->eval_sphere; t->eval_capsule += c->eval_capsule; t->eval_torus += c->eval_torus;
    t->eval_triangle_edge += c->eval_triangle_edge; t->eval_triangle_face += c->eval_triangle_face;
    t->union_lookups += c->union_lookups; t->lower_bound_tests += c->lower_bound_tests;
    t->bound_tests += c->bound_tests; t->intersect_tests += c->intersect_tests;
    if (c->max_steps > t->max_steps) t->max_steps = c->max_steps;
    t->outside_lookups += c->outside_lookups; t->outside_bound_tests += c->outside_bound_tests;
    t->outside_evals += c->outside_evals;
}

static void store_pixel(job *j, int64_t idx, const opixel *px) {
    if (j->out_rgb) { j->out_rgb[3 * idx] = px->color.x; j->out_rgb[3 * idx + 1] = px->color.y; j->out_rgb[3 * idx + 2] = px->color.z; }
    if (j->out_leaf) j->out_leaf[idx] = px->leaf;
    if (j->out_steps) j->out_steps[idx] = (int32_t)px->steps;
    if (j->out_shadow) j->out_shadow[idx] = px->shadow;
}

static void *worker(void *arg) {
    job *j = (job *)arg;
    fto_counters local;
    memset(&local, 0, sizeof(local));
    tl_ctr = j->want_counters ? &local : NULL;
    for (;;) {
        int64_t b = __sync_fetch_and_add(&j->next, j->chunk);
        if (b >= j->n) break;
        int64_t e = b + j->chunk < j->n ? b + j->chunk : j->n;
        for (int64_t i = b; i < e; i++) {
            if (j->mode == 0) {
                /* Array2D.fs:30-38: columns x in parallel, rows y sequential; FColor[X,Y] is x-major */
                int x = j->x0 + (int)i * j->xstride;
                for (int y = 0; y < j->H; y++) {
                    opixel px = scene_trace(j->s, pixel_ray(j->cam, x, y, j->W, j->H, j->eps, j->len));
                    store_pixel(j, (int64_t)x * j->H + y, &px);
                }
            } else if (j->mode == 1) {
                const ft_ray *r = &j->rays[i];
                oray ry;
                ry.origin = V(r->origin[0], r->origin[1], r->origin[2]);
                ry.direction = V(r->direction[0], r->direction[1], r->direction[2]);
                ry.length = r->length; ry.epsilon = r->epsilon;
                opixel px = scene_trace(j->s, ry);
                store_pixel(j, i, &px);
            } else {
                j->out_dist[i] = form_distance(j->s->root->form, V(j->points[3 * i], j->points[3 * i + 1], j->points[3 * i + 2]));
            }
        }
    }
    tl_ctr = NULL;
    if (j->want_counters) {
        pthread_mutex_lock(&j->mu);
        merge_counters(&j->total, &local);
        pthread_mutex_unlock(&j->mu);
    }
    return NULL;
}

static void run_job(job *j, int threads, fto_counters *ctr) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    j->next = j->start;
    j->want_counters = ctr != NULL;
    memset(&j->total, 0, sizeof(j->total));
    pthread_mutex_init(&j->mu, NULL);
    if (threads == 1) {
        worker(j);
    } else {
        pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, worker, j);
        for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
        free(t);
    }
    pthread_mutex_destroy(&j->mu);
    if (ctr) *ctr = j->total;
}

/* Image.render (Image.fs:26-35), columns x_begin, x_begin+x_stride, ... < x_end only; outputs are
 * full-size x-major FColor[X,Y] arrays (only the requested columns written). */
FTO_API int fto_render(const fto_scene *s, const ft_camera *cam, int W, int H, float eps, float len,
                       int x_begin, int x_end, int x_stride, int threads, float *out_rgb, int32_t *out_leaf,
                       int32_t *out_steps, uint32_t *out_shadow, fto_counters *ctr) {
    if (!s || !cam || W <= 0 || H <= 0 || x_begin < 0 || x_end > W || x_stride < 1) return FT_ERR_INVALID_ARGUMENT;
    job j;
    memset(&j, 0, sizeof(j));
    j.s = s; j.cam = cam; j.W = W; j.H = H; j.eps = eps; j.len = len;
    j.out_rgb = out_rgb; j.out_leaf = out_leaf; j.out_steps = out_steps; j.out_shadow = out_shadow;
    j.mode = 0; j.chunk = 1;
    j.start = 0; j.x0 = x_begin; j.xstride = x_stride;
    j.n = x_end > x_begin ? (x_end - x_begin + x_stride - 1) / x_stride : 0;
    run_job(&j, threads, ctr);
    return FT_OK;
}

/* batched SdfScene.trace */
FTO_API int fto_trace_rays(const fto_scene *s, int64_t n, const ft_ray *rays, int threads, float *out_rgb,
                           int32_t *out_leaf, int32_t *out_steps, uint32_t *out_shadow, fto_counters *ctr) {
    if (!s || (n > 0 && !rays)) return FT_ERR_INVALID_ARGUMENT;
    job j;
    memset(&j, 0, sizeof(j));
    j.s = s; j.n = n; j.rays = rays;
    j.out_rgb = out_rgb; j.out_leaf = out_leaf; j.out_steps = out_steps; j.out_shadow = out_shadow;
    j.mode = 1; j.chunk = 256;
    run_job(&j, threads, ctr);
    return FT_OK;
}

/* scene.Object.Form.Distance for n points */
FTO_API int fto_distance(const fto_scene *s, int64_t n, const float *points, int threads, float *out, fto_counters *ctr) {
    if (!s || (n > 0 && (!points || !out))) return FT_ERR_INVALID_ARGUMENT;
    job j;
    memset(&j, 0, sizeof(j));
    j.s = s; j.n = n; j.points = points; j.out_dist = out;
    j.mode = 2; j.chunk = 1024;
    run_job(&j, threads, ctr);
    return FT_OK;
}

/* Distance and Boundary of one described form (KATs for primitives/composites) */
FTO_API int fto_form_distance(const fto_scene *s, int form_id, int64_t n, const float *points, float *out) {
    if (!s || form_id < 0 || form_id >= s->n_forms || !s->forms[form_id]) return FT_ERR_INVALID_ARGUMENT;
    for (int64_t i = 0; i < n; i++)
        out[i] = form_distance(s->forms[form_id], V(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
    return FT_OK;
}
FTO_API int fto_form_boundary(const fto_scene *s, int form_id, float out[4]) {
    if (!s || form_id < 0 || form_id >= s->n_forms || !s->forms[form_id]) return FT_ERR_INVALID_ARGUMENT;
    bnd b = s->forms[form_id]->boundary;
    out[0] = b.c.x; out[1] = b.c.y; out[2] = b.c.z; out[3] = b.r;
    return FT_OK;
}
FTO_API int fto_root_boundary(const fto_scene *s, float out[4]) {
    bnd b = s->root->form->boundary;
    out[0] = b.c.x; out[1] = b.c.y; out[2] = b.c.z; out[3] = b.r;
    return FT_OK;
}

/* ------------------------------------------------------------------------ */
/* Post-process: Image.toColors + Image.toBitmap (Image.fs:37-72, FColor.fs:43-58) */
/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. SC'11): the deterministic replacement, shared with the product by
 * specification, for the reference's racy shared System.Random dither (Image.fs:45-50).          */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
FTO_API void fto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof(c));
}
static unsigned char to_byte(float c, float inv_gamma, float mx, uint32_t bits) {
    float v = powf(c / mx, inv_gamma);                             /* FColor.fs:50-55 MathF.Pow */
    float u = (float)((double)bits * (1.0 / 4294967296.0));        /* range_01 */
    float r = rintf(v * 254.5f + u);                               /* FColor.fs:44 MathF.Round (half to even) */
    int k = (int)r;
    k = k < 255 ? k : 255;
    return (unsigned char)(k < 0 ? 0 : k);
}
/* rgb: FColor[X,Y] x-major.  out: top-down B,G,R scanlines; scanline r, column c = image[X-1-c, r]
 * (Image.fs:66-72: y flip, then Array.rev).                                                       */
FTO_API float fto_to_bitmap(const float *rgb, int X, int Y, float gamma, uint64_t seed, unsigned char *out, int stride) {
    float mx = rgb[0];
    for (int64_t i = 1; i < (int64_t)X * Y * 3; i++) if (rgb[i] > mx) mx = rgb[i]; /* Image.fs:40-42 */
    float raw = mx;
    mx = net_max(0.01f, mx);                                                          /* Image.fs:43 */
    float inv_gamma = 1.0f / gamma;                                                   /* Image.fs:38 */
    for (int x = 0; x < X; x++)
        for (int y = 0; y < Y; y++) {
            int64_t idx = (int64_t)x * Y + y;
            uint32_t c[4] = {(uint32_t)idx, (uint32_t)((uint64_t)idx >> 32), 0u, 0u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            unsigned char *p = out + (int64_t)y * stride + 3 * (int64_t)(X - 1 - x);
            p[2] = to_byte(rgb[3 * idx], inv_gamma, mx, c[0]);
            p[1] = to_byte(rgb[3 * idx + 1], inv_gamma, mx, c[1]);
            p[0] = to_byte(rgb[3 * idx + 2], inv_gamma, mx, c[2]);
        }
    return raw;
}

/* ------------------------------------------------------------------------ */
/* EXTENSION: spectral / refractive ray-test.  NOT reference behaviour: the reference has no   */
/* wavelengths, glass or material-spawned rays.  This is the CPU restatement of this repository's own spec,    */
/* anchored in the reference's dead code Light.fresnel (Light.fs:30-59) and Light.Colors (Light.fs:12-26).     */
/* ------------------------------------------------------------------------ */
static void band_weight(float nu, float w[3]) { /* Light.fs:21-26, frequencies in THz */
    if (nu < 484.0f) { w[0] = 1.0f; w[1] = 0.0f; w[2] = 0.0f; }        /* red    */
    else if (nu < 508.0f) { w[0] = 1.0f; w[1] = 0.5f; w[2] = 0.0f; }   /* orange */
    else if (nu < 526.0f) { w[0] = 1.0f; w[1] = 1.0f; w[2] = 0.0f; }   /* yellow */
    else if (nu < 606.0f) { w[0] = 0.0f; w[1] = 1.0f; w[2] = 0.0f; }   /* green  */
    else if (nu < 668.0f) { w[0] = 0.0f; w[1] = 0.0f; w[2] = 1.0f; }   /* blue   */
    else { w[0] = 0.5f; w[1] = 0.0f; w[2] = 1.0f; }                    /* violet */
}
static float spec(v3 rgb, float nu) {
    float w[3];
    band_weight(nu, w);
    return ((rgb.x * w[0] + rgb.y * w[1]) + rgb.z * w[2]) / ((w[0] + w[1]) + w[2]);
}
static float wavelength_nu(int k, int n) { return 400.0f + 389.0f * (((float)k + 0.5f) / (float)n); }

/* The spectral march.  Primary and shadow rays (armed = 1) are SdfForm.tryTrace as the reference writes it.
 * Continuation rays (armed = 0) start 2 epsilon off the surface they come from: while |Distance| < epsilon they step by
 * epsilon; the first evaluation with |Distance| >= epsilon arms the ray and fixes its medium -- inside glass
 * (Distance < 0: the march continues on -Distance) or outside.  *inside_out receives the medium. */
static int march_signed(const fto_scene *s, oray ray, int armed, oray *hit, int *inside_out) {
    int64_t steps = 0;
    float sign = 1.0f;
    for (;;) {
        if (ray.length <= 0.0f) break;
        if (steps >= s->step_cap) break;
        float D = form_distance(s->root->form, ray.origin);
        steps++;
        if (armed) {
            float d = sign * D;
            if (d < ray.epsilon) { *hit = ray; *inside_out = sign < 0.0f; return 1; }
            ray = ray_move(ray, d);
        } else {
            sign = D < 0.0f ? -1.0f : 1.0f;
            float d = sign * D;
            if (d >= ray.epsilon) { armed = 1; ray = ray_move(ray, d); }
            else ray = ray_move(ray, ray.epsilon);
        }
    }
    *inside_out = sign < 0.0f;
    return 0;
}

typedef struct { int n; float lcos[FT_SPECTRAL_MAX_LIGHTS]; v3 intensity[FT_SPECTRAL_MAX_LIGHTS]; } lit_set;

/* which lights reach the hit (the reference's light loop, SdfScene.fs:13-26), wavelength independent */
static void gather_lights(const fto_scene *s, const oray *hit, v3 normal, int inside, lit_set *ls) {
    oray moved = ray_move(*hit, -hit->epsilon); /* SdfObject.fs:73 */
    v3 pos = moved.origin;
    ls->n = 0;
    for (int l = 0; l < s->n_lights; l++) {
        const olight *lt = &s->lights[l];
        v3 ldir = (lt->kind == FT_LIGHT_DIRECTIONAL) ? lt->v : vnormalize(vsub(lt->v, pos));
        float lcos = vdot(normal, ldir);
        if (!(lcos > 0.0f)) continue;
        oray sr;
        v3 intensity;
        sr.origin = pos;
        sr.epsilon = hit->epsilon;
        if (lt->kind == FT_LIGHT_DIRECTIONAL) { sr.direction = lt->v; sr.length = 1000.0f; intensity = lt->color; }
        else {
            v3 diff = vsub(lt->v, pos);
            float d2 = vlen2(diff);
            sr.direction = vdiv(diff, d2);
            sr.length = sqrtf(d2);
            intensity = vdiv(lt->color, d2);
        }
        int shadowed = 0;
        if (!inside) { oray dummy; int dm; shadowed = march_signed(s, sr, 1, &dummy, &dm); }
        if (!shadowed) { ls->lcos[ls->n] = lcos; ls->intensity[ls->n] = intensity; ls->n++; }
    }
}
static float shade_solid(const fto_scene *s, v3 albedo, const lit_set *ls, float nu) {
    float lc = spec(s->background, nu); /* SdfScene.fs:12 */
    for (int i = 0; i < ls->n; i++) lc = lc + spec(ls->intensity[i], nu) * ls->lcos[i];
    const float pi_inv = 1.0f / 3.14159274f;
    return spec(albedo, nu) * (lc * pi_inv); /* SdfScene.fs:28 */
}

/* Light.fresnel n1 n2 normal direction (Light.fs:30-59) as written + the choice of the continuation */
static void glass_interact(const fto_scene *s, v3 ab, float nu, uint64_t seed, uint32_t pixel, uint32_t k, uint32_t bounce,
                           const oray *hit, v3 normal, int inside, oray *next) {
    (void)s;
    float lam = 299.792458f / nu;
    float nn = ab.x + ab.y / (lam * lam);
    float n1 = inside ? nn : 1.0f, n2 = inside ? 1.0f : nn;
    v3 d = hit->direction;
    v3 N = vdot(normal, d) > 0.0f ? normal : vneg(normal);
    float n1n2 = n1 / n2;
    float cosi = vdot(N, d);
    float n1n2cosi = n1n2 * cosi;
    float cost = sqrtf(1.0f + n1n2cosi * cosi - n1n2);
    float reflectance = 1.0f;
    if (!isnan(cost)) {
        float a = n2 * cosi, b = n1 * cost;
        float x = (a - b) / (a + b);
        float rs = x * x;
        a = n1 * cosi; b = n2 * cost;
        x = (a - b) / (a + b);
        float rp = x * x;
        reflectance = 0.5f * (rs + rp);
    }
    uint32_t c[4] = {pixel, k, bounce, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float u = (float)((double)c[0] * (1.0 / 4294967296.0));
    v3 nd;
    if (u < reflectance) nd = vnormalize(vsub(d, vscale(N, 2.0f * cosi)));
    else nd = vnormalize(vadd(vscale(d, n1n2), vscale(N, n1n2cosi - cost)));
    next->origin = vadd(hit->origin, vscale(nd, 2.0f * hit->epsilon));
    next->direction = nd;
    next->length = hit->length - 2.0f * hit->epsilon;
    next->epsilon = hit->epsilon;
}

typedef struct {
    const fto_scene *s; const ft_camera *cam; int W, H; float eps, len; int nl, max_bounces; uint64_t seed;
    float *out; volatile int64_t next;
    int x0, xstride, ncols; /* columns x0, x0 + xstride, ... (ncols of them) */
} sjob;

static void spectral_pixel(const sjob *j, int x, int y, float *rgb) {
    const fto_scene *s = j->s;
    const int nl = j->nl;
    uint32_t pixel = (uint32_t)((int64_t)x * j->H + y);
    float L[64];
    oray ray = pixel_ray(j->cam, x, y, j->W, j->H, j->eps, j->len), hit;
    /* wave 0: one primary ray for all wavelengths */
    int inside0;
    if (!march_signed(s, ray, 1, &hit, &inside0)) {
        for (int k = 0; k < nl; k++) L[k] = spec(s->background, wavelength_nu(k, nl));
    } else {
        v3 normal = form_normal_from_ray(s->root->form, &hit);
        const oobj *leaf = object_material_leaf(s->root, hit.origin);
        if (s->material_kind[leaf->material] != FT_MATERIAL_GLASS) {
            lit_set ls;
            gather_lights(s, &hit, normal, 0, &ls);
            for (int k = 0; k < nl; k++) L[k] = shade_solid(s, s->material_rgb[leaf->material], &ls, wavelength_nu(k, nl));
        } else {
            for (int k = 0; k < nl; k++) { /* one continuation path per wavelength */
                float nu = wavelength_nu(k, nl);
                int inside = 0;
                oray cur_hit = hit;
                v3 cur_normal = normal;
                v3 ab = s->material_rgb[leaf->material];
                float val = 0.0f;
                for (uint32_t bounce = 0;; bounce++) {
                    if ((int)bounce >= j->max_bounces) { val = 0.0f; break; }
                    oray nxt;
                    glass_interact(s, ab, nu, j->seed, pixel, (uint32_t)k, bounce, &cur_hit, cur_normal, inside, &nxt);
                    if (!march_signed(s, nxt, 0, &cur_hit, &inside)) { val = spec(s->background, nu); break; }
                    cur_normal = form_normal_from_ray(s->root->form, &cur_hit);
                    const oobj *lf = object_material_leaf(s->root, cur_hit.origin);
                    if (s->material_kind[lf->material] != FT_MATERIAL_GLASS) {
                        lit_set ls;
                        gather_lights(s, &cur_hit, cur_normal, inside, &ls);
                        val = shade_solid(s, s->material_rgb[lf->material], &ls, nu);
                        break;
                    }
                    ab = s->material_rgb[lf->material];
                }
                L[k] = val;
            }
        }
    }
    float acc[3] = {0, 0, 0}, den[3] = {0, 0, 0};
    for (int k = 0; k < nl; k++) {
        float w[3];
        band_weight(wavelength_nu(k, nl), w);
        for (int c = 0; c < 3; c++) { acc[c] = acc[c] + w[c] * L[k]; den[c] = den[c] + w[c]; }
    }
    for (int c = 0; c < 3; c++) rgb[c] = den[c] > 0.0f ? acc[c] / den[c] : 0.0f;
}
static void *spectral_worker(void *arg) {
    sjob *j = (sjob *)arg;
    for (;;) {
        int64_t i = __sync_fetch_and_add(&j->next, 1);
        if (i >= j->ncols) break;
        int64_t x = j->x0 + i * j->xstride;
        for (int y = 0; y < j->H; y++) spectral_pixel(j, (int)x, y, j->out + 3 * ((int64_t)x * j->H + y));
    }
    return NULL;
}
/* columns x_begin, x_begin + x_stride, ... < x_end only (the output is the full-size x-major array) */
FTO_API int fto_render_spectral(const fto_scene *s, const ft_camera *cam, int W, int H, float eps, float len, int n_wavelengths,
                                int max_bounces, uint64_t seed, int x_begin, int x_end, int x_stride, int threads,
                                float *out_rgb) {
    if (!s || !cam || !out_rgb || W <= 0 || H <= 0 || n_wavelengths < 1 || n_wavelengths > 64 || max_bounces < 0 ||
        s->n_lights > FT_SPECTRAL_MAX_LIGHTS || x_begin < 0 || x_end > W || x_stride < 1)
        return FT_ERR_INVALID_ARGUMENT;
    sjob j;
    memset(&j, 0, sizeof(j));
    j.s = s; j.cam = cam; j.W = W; j.H = H; j.eps = eps; j.len = len; j.nl = n_wavelengths; j.max_bounces = max_bounces;
    j.seed = seed; j.out = out_rgb; j.next = 0;
    j.x0 = x_begin; j.xstride = x_stride; j.ncols = x_end > x_begin ? (x_end - x_begin + x_stride - 1) / x_stride : 0;
    if (threads < 1) threads = 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, spectral_worker, &j);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
    return FT_OK;
}

/* ------------------------------------------------------------------------ */
/* EXTENSION: volumetric cloud material.  NOT reference behaviour: the reference has no media,    */
/* no stochastic sampling and no anti-aliasing.  CPU restatement of this repository's own spec; the density      */
/* field is the reference's dead-code Noise.gradient (Noise.fs:78-110) restated as written, float64.             */
/* ------------------------------------------------------------------------ */
typedef struct { const int32_t *perm; const double *grad; } noise_tab;
static inline int nperm(const noise_tab *n, int x) { return n->perm[x & 255]; }                       /* Noise.fs:25-26 */
static inline int nindex(const noise_tab *n, int ix, int iy, int iz) {                                /* Noise.fs:28-29 */
    return nperm(n, ix + nperm(n, iy + nperm(n, iz)));
}
static inline double glattice(const noise_tab *n, int ix, int iy, int iz, double fx, double fy, double fz) {
    const double *g = n->grad + 3 * nindex(n, ix, iy, iz);
    return ((g[0] * fx + g[1] * fy) + g[2]) + fz; /* Noise.fs:80 as written: x*fx + y*fy + z+fz */
}
static inline double nlerp(double t, double x0, double x1) { return x0 + t * (x1 - x0); }             /* Noise.fs:82 */
static inline double smooth_step(double x) { return (x * x) * (3.0 - 2.0 * x); }                      /* Noise.fs:83 */
static inline double lat_and_lerp(const noise_tab *n, double w, int ix0, int iy0, int iz0, double fx0, double fy0, double fz0,
                                  int ix1, int iy1, int iz1, double fx1, double fy1, double fz1) {  /* Noise.fs:92-95 */
    double a = glattice(n, ix0, iy0, iz0, fx0, fy0, fz0);
    double b = glattice(n, ix1, iy1, iz1, fx1, fy1, fz1);
    return nlerp(w, a, b);
}
static double noise_gradient(const noise_tab *n, double x, double y, double z) {                      /* Noise.fs:97-110 */
    int ix = (int)floor(x), iy = (int)floor(y), iz = (int)floor(z);                                   /* Noise.fs:85-90 */
    double fx0 = x - (double)ix, fy0 = y - (double)iy, fz0 = z - (double)iz;
    double fx1 = fx0 - 1.0, fy1 = fy0 - 1.0, fz1 = fz0 - 1.0;
    double wx = smooth_step(fx0), wy = smooth_step(fy0), wz = smooth_step(fz0);
    double vy0 = lat_and_lerp(n, wx, ix, iy, iz, fx0, fy0, fz0, ix + 1, iy, iz, fx1, fy0, fz0);
    double vy1 = lat_and_lerp(n, wx, ix, iy + 1, iz, fx0, fy1, fz0, ix + 1, iy + 1, iz, fx1, fy1, fz0);
    double vz0 = nlerp(wy, vy0, vy1);
    vy0 = lat_and_lerp(n, wx, ix, iy, iz + 1, fx0, fy0, fz1, ix + 1, iy, iz + 1, fx1, fy0, fz1);
    vy1 = lat_and_lerp(n, wx, ix, iy + 1, iz + 1, fx0, fy1, fz1, ix + 1, iy + 1, iz + 1, fx1, fy1, fz1);
    double vz1 = nlerp(wy, vy0, vy1);
    return nlerp(wz, vz0, vz1);
}
FTO_API double fto_noise_gradient(const int32_t *perm, const double *gradients, double x, double y, double z) {
    noise_tab n = {perm, gradients};
    return noise_gradient(&n, x, y, z);
}

typedef struct {
    const fto_scene *s; const ft_camera *cam; int W, H; float eps, len; const ft_cloud_params *cp; int spp; uint64_t seed;
    float *out; volatile int64_t next;
    int64_t secondary, track_steps, scattered; pthread_mutex_t mu;
    int x0, xstride, ncols;
} cjob;
typedef struct { int64_t secondary, track_steps, scattered; } cstat;

static inline float clamp01f(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

/* local collision rate: density * clamp01(bias + gain * noise(p f)) * clamp01(-Distance / fade) */
static float cloud_density(const ft_cloud_params *cp, v3 p, float D) {
    noise_tab n = {cp->perm, cp->gradients};
    double f = (double)cp->frequency;
    float nz = (float)noise_gradient(&n, (double)p.x * f, (double)p.y * f, (double)p.z * f);
    float shape = clamp01f(cp->bias + cp->gain * nz);
    float edge = clamp01f(-D / cp->fade);
    return (cp->density * shape) * edge;
}
/* Random numbers of the extension: block b of stream `stream` of (pixel, sample) is
 * Philox4x32-10(counter = {pixel, sample, b, stream}, key = seed), four 32-bit words.  Stream 0 / block 0 holds the
 * sub-pixel jitter (words 0, 1).  A tracking stream uses block 0 word 0 for the collision draw of step 0 and word 1 for
 * the offset xi0 of the first step; the collision draw of step k >= 1 is word (k + 1) & 3 of block (k + 1) >> 2, so
 * that one Philox evaluation serves four tracking steps. */
static float cloud_u(uint32_t word) { return (float)((double)word * (1.0 / 4294967296.0)); }
static void cloud_block(const cjob *j, uint32_t pixel, uint32_t sample, uint32_t block, uint32_t stream, uint32_t c[4]) {
    c[0] = pixel; c[1] = sample; c[2] = block; c[3] = stream;
    philox4x32_10(c, (uint32_t)j->seed, (uint32_t)(j->seed >> 32));
}
static void cloud_rnd(const cjob *j, uint32_t pixel, uint32_t sample, uint32_t step, uint32_t stream, float u[2]) {
    uint32_t c[4];
    cloud_block(j, pixel, sample, step, stream, c);
    u[0] = cloud_u(c[0]);
    u[1] = cloud_u(c[1]);
}
/* collision draw of tracking step k >= 1; `blk` holds the stream's current block (block 0 after the stream's first draw) */
static float cloud_track_u(const cjob *j, uint32_t pixel, uint32_t sample, uint32_t k, uint32_t stream, uint32_t blk[4]) {
    const uint32_t jj = k + 1u;
    if ((jj & 3u) == 0u) cloud_block(j, pixel, sample, jj >> 2, stream, blk);
    return cloud_u(blk[jj & 3u]);
}
/* the reference's shadow ray of light lt from `origin` (SdfLight.fs:10-20 / 26-41): 1 when unoccluded */
static int shadow_unoccluded(const fto_scene *s, const olight *lt, v3 origin, float eps) {
    oray sr;
    sr.origin = origin;
    sr.epsilon = eps;
    if (lt->kind == FT_LIGHT_DIRECTIONAL) { sr.direction = lt->v; sr.length = 1000.0f; }
    else {
        v3 diff = vsub(lt->v, origin);
        float d2 = vlen2(diff);
        sr.direction = vdiv(diff, d2);
        sr.length = sqrtf(d2);
    }
    oray dummy;
    return !form_try_trace(s, s->root->form, sr, &dummy, NULL);
}
/* visibility of light lt from the scatter point p: fixed-step tracking through the medium, then the shadow ray */
static int light_track(const cjob *j, cstat *st, const olight *lt, v3 p, v3 ldir, float maxdist, uint32_t pixel, uint32_t sample,
                       uint32_t stream) {
    const ft_cloud_params *cp = j->cp;
    float u[2];
    uint32_t blk[4];
    cloud_block(j, pixel, sample, 0u, stream, blk);
    u[0] = cloud_u(blk[0]);
    const float xi0 = cloud_u(blk[1]);
    st->secondary++;
    for (uint32_t k = 0;; k++) {
        float t = ((float)k + xi0) * cp->step;
        if (t >= maxdist) return 1;
        if (k >= FT_CLOUD_MAX_TRACK_STEPS) return 0;
        v3 q = vadd(p, vscale(ldir, t));
        float D = form_distance(j->s->root->form, q);
        st->track_steps++;
        if (D >= j->eps) { st->secondary++; return shadow_unoccluded(j->s, lt, q, j->eps); }
        if (k > 0) u[0] = cloud_track_u(j, pixel, sample, k, stream, blk);
        if (u[0] < cloud_density(cp, q, D) * cp->step) return 0;
    }
}

static v3 cloud_sample(const cjob *j, cstat *st, int x, int y, uint32_t sample) {
    const fto_scene *s = j->s;
    const ft_cloud_params *cp = j->cp;
    const uint32_t pixel = (uint32_t)((int64_t)x * j->H + y);
    const uint32_t nstreams = 1u + (uint32_t)s->n_lights;
    float u[2];
    cloud_rnd(j, pixel, sample, 0u, 0u, u); /* sub-pixel jitter */
    float max_size = (float)(j->W > j->H ? j->W : j->H);
    float pu = ((float)x + u[0]) / max_size, pv = ((float)y + u[1]) / max_size;
    v3 fwd = V(j->cam->forward[0], j->cam->forward[1], j->cam->forward[2]);
    v3 rs = V(j->cam->right_scaled[0], j->cam->right_scaled[1], j->cam->right_scaled[2]);
    v3 us = V(j->cam->up_scaled[0], j->cam->up_scaled[1], j->cam->up_scaled[2]);
    oray ray;
    ray.origin = V(j->cam->position[0], j->cam->position[1], j->cam->position[2]);
    ray.direction = vnormalize(vadd(vadd(fwd, vscale(rs, pu - 0.5f)), vscale(us, pv - 0.5f)));
    ray.epsilon = j->eps;
    ray.length = j->len;
    const float pi_inv = 1.0f / 3.14159274f;
    for (uint32_t seg = 0;; seg++) {
        if (seg >= FT_CLOUD_MAX_SEGMENTS) return s->background;
        if (seg > 0) st->secondary++;
        oray hit;
        if (!form_try_trace(s, s->root->form, ray, &hit, NULL)) return s->background; /* SdfScene.fs:10 */
        v3 normal = form_normal_from_ray(s->root->form, &hit);
        const oobj *leaf = object_material_leaf(s->root, hit.origin);
        if (s->material_kind[leaf->material] != FT_MATERIAL_CLOUD) { /* the reference's shading, SdfScene.fs:12-28 */
            v3 pos = ray_move(hit, -hit.epsilon).origin;
            v3 lc = s->background;
            for (int l = 0; l < s->n_lights; l++) {
                const olight *lt = &s->lights[l];
                v3 ldir = (lt->kind == FT_LIGHT_DIRECTIONAL) ? lt->v : vnormalize(vsub(lt->v, pos));
                float lcos = vdot(normal, ldir);
                if (!(lcos > 0.0f)) continue;
                v3 intensity = lt->color;
                if (lt->kind != FT_LIGHT_DIRECTIONAL) intensity = vdiv(lt->color, vlen2(vsub(lt->v, pos)));
                st->secondary++;
                if (shadow_unoccluded(s, lt, pos, hit.epsilon)) lc = vadd(lc, vscale(intensity, lcos));
            }
            return vmul(s->material_rgb[leaf->material], vscale(lc, pi_inv));
        }
        /* cloud: fixed-step tracking along the camera ray from the hit point */
        const uint32_t stream = 1u + seg * nstreams;
        uint32_t blk[4];
        cloud_block(j, pixel, sample, 0u, stream, blk);
        u[0] = cloud_u(blk[0]);
        const float xi0 = cloud_u(blk[1]);
        int exited = 0;
        for (uint32_t k = 0;; k++) {
            float t = ((float)k + xi0) * cp->step;
            if (hit.length - t <= 0.0f || k >= FT_CLOUD_MAX_TRACK_STEPS) return s->background;
            v3 p = vadd(hit.origin, vscale(hit.direction, t));
            float D = form_distance(s->root->form, p);
            st->track_steps++;
            if (D >= hit.epsilon) { /* left the medium: the camera ray goes on as an ordinary ray */
                ray.origin = p;
                ray.length = hit.length - t;
                exited = 1;
                break;
            }
            if (k > 0) u[0] = cloud_track_u(j, pixel, sample, k, stream, blk);
            if (u[0] < cloud_density(cp, p, D) * cp->step) { /* real collision: single scattering */
                st->scattered++;
                v3 lc = s->background;
                for (int l = 0; l < s->n_lights; l++) {
                    const olight *lt = &s->lights[l];
                    v3 ldir, intensity = lt->color;
                    float maxdist = INFINITY;
                    if (lt->kind == FT_LIGHT_DIRECTIONAL) ldir = lt->v;
                    else {
                        v3 diff = vsub(lt->v, p);
                        float d2 = vlen2(diff);
                        maxdist = sqrtf(d2);
                        ldir = vdiv(diff, maxdist);
                        intensity = vdiv(lt->color, d2);
                    }
                    if (light_track(j, st, lt, p, ldir, maxdist, pixel, sample, stream + 1u + (uint32_t)l)) lc = vadd(lc, intensity);
                }
                return vmul(s->material_rgb[leaf->material], vscale(lc, pi_inv * 0.25f)); /* isotropic phase 1/(4 pi) */
            }
        }
        if (!exited) return s->background;
    }
}
static void *cloud_worker(void *arg) {
    cjob *j = (cjob *)arg;
    cstat st = {0, 0, 0};
    for (;;) {
        int64_t i = __sync_fetch_and_add(&j->next, 1);
        if (i >= j->ncols) break;
        int64_t x = j->x0 + i * j->xstride;
        for (int y = 0; y < j->H; y++) {
            v3 sum = V(0.0f, 0.0f, 0.0f);
            for (int g = 0; g < FT_CLOUD_SAMPLE_GROUPS; g++) { /* group sums in sample order, added in group order */
                int s0 = (int)(((int64_t)g * j->spp) / FT_CLOUD_SAMPLE_GROUPS), s1 = (int)(((int64_t)(g + 1) * j->spp) / FT_CLOUD_SAMPLE_GROUPS);
                v3 part = V(0.0f, 0.0f, 0.0f);
                for (int k = s0; k < s1; k++) part = vadd(part, cloud_sample(j, &st, (int)x, y, (uint32_t)k));
                sum = g == 0 ? part : vadd(sum, part);
            }
            float *o = j->out + 3 * ((int64_t)x * j->H + y);
            o[0] = sum.x / (float)j->spp; o[1] = sum.y / (float)j->spp; o[2] = sum.z / (float)j->spp;
        }
    }
    pthread_mutex_lock(&j->mu);
    j->secondary += st.secondary; j->track_steps += st.track_steps; j->scattered += st.scattered;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}
/* counters (optional): {secondary marches, tracking steps, scattered samples} */
FTO_API int fto_render_cloud(const fto_scene *s, const ft_camera *cam, int W, int H, float eps, float len, const ft_cloud_params *cp,
                             int spp, uint64_t seed, int x_begin, int x_end, int x_stride, int threads, float *out_rgb,
                             int64_t *counters) {
    if (!s || !cam || !cp || !cp->perm || !cp->gradients || !out_rgb || W <= 0 || H <= 0 || spp < 1 || !(cp->step > 0.0f) ||
        !(cp->fade > 0.0f) || x_begin < 0 || x_end > W || x_stride < 1)
        return FT_ERR_INVALID_ARGUMENT;
    cjob j;
    memset(&j, 0, sizeof(j));
    j.s = s; j.cam = cam; j.W = W; j.H = H; j.eps = eps; j.len = len; j.cp = cp; j.spp = spp; j.seed = seed; j.out = out_rgb;
    j.x0 = x_begin; j.xstride = x_stride; j.ncols = x_end > x_begin ? (x_end - x_begin + x_stride - 1) / x_stride : 0;
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, cloud_worker, &j);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
    pthread_mutex_destroy(&j.mu);
    if (counters) { counters[0] = j.secondary; counters[1] = j.track_steps; counters[2] = j.scattered; }
    return FT_OK;
}
