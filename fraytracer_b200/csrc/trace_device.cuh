// trace_device.cuh -- device-side restatement of the reference's per-ray path over the compiled scene.
// Included by kernels_strict.cu (-fmad=false: every float32 operation individually rounded, bit-identical
// to the reference's arithmetic) and kernels_fast.cu (-fmad=true), each inside its own namespace FT_NS.
//
// Reference paths are relative to /root/reference/src/FrayTracer.
#pragma once
#include <cuda_runtime.h>

#include "device_scene.h"
#include "ft_math.h"

namespace FT_NS {
using namespace ftm;

// algorithmic float32 operation counts per node evaluation (DESIGN.md "flop accounting")
constexpr int FLOPS_SPHERE = 10;
constexpr int FLOPS_CAPSULE = 21;
constexpr int FLOPS_TORUS = 27;
constexpr int FLOPS_TRI_EDGE = 85;
constexpr int FLOPS_TRI_FACE = 34;
constexpr int FLOPS_LB_TEST = 2;      // lb - dtc, compare
constexpr int FLOPS_BOUND_TEST = 11;  // distance (9) - r, compare
constexpr int FLOPS_CELL_LOOKUP = 19; // 3 sub, 3 mul, 3 floor, distance to centre (9) + 1
constexpr int FLOPS_MOVE = 7;         // Ray.move
constexpr int FLOPS_MINMAX = 1;
constexpr int FLOPS_SMOOTH_TERM = 3;  // mul, exp, add
constexpr int FLOPS_NORMAL = 12 + 7;  // 3 sub + normalize (9) + Ray.get
constexpr int FLOPS_LIGHT = 16;       // dot (5) + compare + scale-add (6) [+ point light extras counted apart]
constexpr int FLOPS_POINT_LIGHT = 24; // normalize(pos-p) 12, diff 3, len2 5, sqrt, 3 div
constexpr int FLOPS_FINAL = 6;
constexpr int FLOPS_PRIMARY_RAY = 25; // Camera.uniformPixelToRay

struct Acc {
    unsigned long long flops;
    unsigned long long steps;
};

template <bool COUNT>
__device__ __forceinline__ void add_flops(Acc &a, int n) {
    if (COUNT) a.flops += (unsigned)n;
}

__device__ __forceinline__ float4 ld4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ V3 xyz(float4 v) { return V3{v.x, v.y, v.z}; }

// ---- primitives (SdfForm.fs:117-268); b0 = record[0] = bounding sphere, already loaded
template <bool COUNT>
__device__ __forceinline__ float prim_distance(const DevScene &S, uint32_t word, float4 b0, V3 p, Acc &acc) {
    const uint32_t kind = FT_WORD_KIND(word);
    const float4 *rec = S.nodes + FT_WORD_OFF(word);
    if (kind == FT_K_SPHERE) {  // :127-129  Vector3.Distance(Center, position) - Radius
        add_flops<COUNT>(acc, FLOPS_SPHERE);
        return distance(xyz(b0), p) - b0.w;
    }
    if (kind == FT_K_TORUS) {  // :189-194
        add_flops<COUNT>(acc, FLOPS_TORUS);
        const float4 r1 = ld4(rec + 1), r2 = ld4(rec + 2);
        const V3 n = xyz(r1);
        const float dplane = dot(p, n) + r1.w;
        const float dcenter = distance(xyz(b0), p - n * dplane);
        const float dcircle = dcenter - r2.x;
        return sqrtf(dplane * dplane + dcircle * dcircle) - r2.y;  // Vector2.Length
    }
    if (kind == FT_K_CAPSULE) {  // :151-164
        add_flops<COUNT>(acc, FLOPS_CAPSULE);
        const float4 r1 = ld4(rec + 1), r2 = ld4(rec + 2), r3 = ld4(rec + 3);
        const V3 diff = p - xyz(r1);
        const float t = dot(diff, xyz(r3));
        float d;
        if (t <= 0.0f) d = length(diff);
        else if (t >= 1.0f) d = distance(diff, xyz(r2));
        else d = distance(diff, xyz(r2) * t);
        return d - r1.w;
    }
    {  // triangle :227-250
        const float4 r1 = ld4(rec + 1);
        const V3 p1 = p - xyz(r1), p2 = p - xyz(ld4(rec + 2)), p3 = p - xyz(ld4(rec + 3));
        const float s1 = dot(xyz(ld4(rec + 11)), p1), s2 = dot(xyz(ld4(rec + 12)), p2), s3 = dot(xyz(ld4(rec + 13)), p3);
        const int sg = ((s1 > 0.f) - (s1 < 0.f)) + ((s2 > 0.f) - (s2 < 0.f)) + ((s3 > 0.f) - (s3 < 0.f));
        float d;
        if (sg < 2) {
            add_flops<COUNT>(acc, FLOPS_TRI_EDGE);
            const V3 v21 = xyz(ld4(rec + 4)), v32 = xyz(ld4(rec + 5)), v13 = xyz(ld4(rec + 6));
            // clamp01 x = Max(0, Min(1, x))
            const float t21 = fmaxf(0.f, fminf(1.f, dot(xyz(ld4(rec + 7)), p1)));
            const float t32 = fmaxf(0.f, fminf(1.f, dot(xyz(ld4(rec + 8)), p2)));
            const float t13 = fmaxf(0.f, fminf(1.f, dot(xyz(ld4(rec + 9)), p3)));
            const float d21 = distance2(p1, v21 * t21);
            const float d32 = distance2(p2, v32 * t32);
            const float d13 = distance2(p3, v13 * t13);
            d = sqrtf(fminf(d13, fminf(d32, d21)));
        } else {
            add_flops<COUNT>(acc, FLOPS_TRI_FACE);
            d = fabsf(dot(xyz(ld4(rec + 10)), p1));
        }
        return d - r1.w;
    }
}

// ---- spatial lookup (SdfBoundary.fs:276-282): returns the cell's first entry index and dist(p, centre)
template <bool COUNT>
__device__ __forceinline__ void grid_cell(const DevScene &S, uint32_t uoff, V3 p, uint32_t &begin, float &dtc, Acc &acc) {
    add_flops<COUNT>(acc, FLOPS_CELL_LOOKUP);
    const float4 g1 = ld4(S.nodes + uoff + 1), g2 = ld4(S.nodes + uoff + 2), g3 = ld4(S.nodes + uoff + 3);
    const V3 c = (p - xyz(g1)) * xyz(g2);
    const int mx = __float_as_int(g3.x), my = __float_as_int(g3.y), mz = __float_as_int(g3.z);
    const int ix = clamp_int(0, mx, floor_to_int(c.x));
    const int iy = clamp_int(0, my, floor_to_int(c.y));
    const int iz = clamp_int(0, mz, floor_to_int(c.z));
    const uint32_t ci = __float_as_uint(g1.w) + (uint32_t)(ix * __float_as_int(g3.w) + iy * (mz + 1) + iz);
    const float4 cell = ld4(S.cells + ci);
    dtc = distance(p, xyz(cell));  // SdfForm.fs:25
    begin = __float_as_uint(cell.w);
}

// one 32-byte candidate record (device_scene.h): two independent LDG.128 from the same sector
struct Entry {
    float4 a;  // {LowerBound, bits(child word), cx, cy}
    float4 b;  // {cz, r, bits(child position), -}
    __device__ __forceinline__ float lb() const { return a.x; }
    __device__ __forceinline__ uint32_t word() const { return __float_as_uint(a.y); }
    __device__ __forceinline__ float4 bound() const { return float4{a.z, a.w, b.x, b.y}; }
    __device__ __forceinline__ int child() const { return __float_as_int(b.z); }
};
__device__ __forceinline__ Entry ld_entry(const DevScene &S, uint32_t i) {
    const float4 *p = S.entries + 2ull * i;
    return Entry{__ldg(p), __ldg(p + 1)};
}

// ---- the second pruning test, SdfForm.fs:31 / SdfObject.fs:38:  min > Vector3.Distance(C, p) - R.
// s = |C - p|^2 is formed exactly like the reference; the root is only taken when the outcome is open:
//   m + R <= 0              =>  sqrt(s) - R >= -R >= m          =>  false
//   s > (m+R)^2 (1 + 2^-12)  =>  fl(sqrt(s)) > m + R (the 2^-13 margin dwarfs the 3 roundings involved)
//                            =>  fl(sqrt(s)) - R > m, and rounding is monotone  =>  false
// Otherwise bd = fl(fl(sqrt(s)) - R) is computed as the reference does and compared.  `bd` is only valid
// when the function returns true.  NaNs fall through to the exact path (all comparisons false).
__device__ __forceinline__ bool bound_test(float m, float4 b0, V3 p, float &bd) {
    const V3 dl = xyz(b0) - p;
    const float s = dot(dl, dl);
    const float u = m + b0.w;
    if (u <= 0.0f || s > (u * u) * 1.000244140625f) return false;
    bd = sqrtf(s) - b0.w;
    return m > bd;
}

// ---- union candidate scan (SdfForm.fs:27-33) from entry i with running minimum m.
// Returns true when the list is exhausted (early exit on the sorted LowerBound, see device_scene.h);
// false when entry i is a composite child that must be evaluated by the caller.
// The next record is fetched before the current one is tested (the list is closed by a sentinel and the
// array by two padding records, so the speculative load is always in bounds).
template <bool COUNT>
__device__ __forceinline__ bool union_scan(const DevScene &S, V3 p, float dtc, uint32_t &i, float &m, uint32_t &child, Acc &acc) {
    Entry e = ld_entry(S, i);
    for (;;) {
        const Entry nx = ld_entry(S, i + 1);
        add_flops<COUNT>(acc, FLOPS_LB_TEST);
        if (!(m > e.lb() - dtc)) return true;
        const float4 b0 = e.bound();
        float bd;
        add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
        if (bound_test(m, b0, p, bd)) {  // m > SdfBoundary.getMinDistance boundary position
            const uint32_t w = e.word();
            if (FT_WORD_KIND(w) == FT_K_SPHERE) {
                m = bd;  // the sphere's Distance is the same expression: Min(m, bd) = bd here
                add_flops<COUNT>(acc, FLOPS_MINMAX);
            } else if (FT_IS_PRIMITIVE(w)) {
                m = fminf(m, prim_distance<COUNT>(S, w, b0, p, acc));
                add_flops<COUNT>(acc, FLOPS_MINMAX);
            } else {
                child = w;
                return false;
            }
        }
        e = nx;
        i++;
    }
}

struct Frame {
    uint32_t word;
    uint32_t idx;
    float acc;
    float aux;
};

// ---- Form.Distance position for an arbitrary compiled form (the closure of Types.fs:42)
template <bool COUNT>
__device__ __forceinline__ float eval_form(const DevScene &S, uint32_t root, V3 p, Acc &acc) {
    if (FT_IS_PRIMITIVE(root)) return prim_distance<COUNT>(S, root, ld4(S.nodes + FT_WORD_OFF(root)), p, acc);
    Frame stack[FT_MAX_DEPTH];
    int sp = 0;
    uint32_t cur = root;
    float val = 0.f;
    for (;;) {
        // ------------------------------------------------ descend: evaluate `cur`
        bool have_val = false;
        while (!have_val) {
            const uint32_t kind = FT_WORD_KIND(cur), off = FT_WORD_OFF(cur);
            if (kind < 8u) {
                val = prim_distance<COUNT>(S, cur, ld4(S.nodes + off), p, acc);
                have_val = true;
            } else if (kind == FT_K_UNION) {  // SdfForm.fs:22-34
                uint32_t i;
                float dtc;
                grid_cell<COUNT>(S, off, p, i, dtc, acc);
                const Entry e0 = ld_entry(S, i);
                const uint32_t first = e0.word();
                if (FT_IS_PRIMITIVE(first)) {
                    float m = prim_distance<COUNT>(S, first, e0.bound(), p, acc);  // :26
                    i++;
                    uint32_t child;
                    if (union_scan<COUNT>(S, p, dtc, i, m, child, acc)) {
                        val = m;
                        have_val = true;
                    } else {
                        stack[sp++] = Frame{cur, i, m, dtc};
                        cur = child;
                    }
                } else {
                    stack[sp++] = Frame{cur, i, __int_as_float(0x7f800000), dtc};
                    cur = first;
                }
            } else if (kind == FT_K_SUBTRACT) {  // SdfForm.fs:45-47
                const float4 h = ld4(S.nodes + off + 1);
                stack[sp++] = Frame{cur, 0u, 0.f, 0.f};
                cur = __float_as_uint(h.x);
            } else if (kind == FT_K_INTERSECT) {  // SdfForm.fs:58-64
                const float4 h = ld4(S.nodes + off + 1);
                stack[sp++] = Frame{cur, 0u, 0.f, 0.f};
                cur = __ldg(S.links + __float_as_uint(h.y));
            } else {  // FT_K_SMOOTH  SdfForm.fs:75-82
                const float4 h = ld4(S.nodes + off + 1);
                stack[sp++] = Frame{cur, 0u, 0.f, 0.f};
                cur = __ldg(S.links + __float_as_uint(h.w));
            }
        }
        // ------------------------------------------------ ascend: feed `val` to the enclosing composites
        bool descend = false;
        while (!descend) {
            if (sp == 0) return val;
            Frame &f = stack[sp - 1];
            const uint32_t kind = FT_WORD_KIND(f.word), off = FT_WORD_OFF(f.word);
            if (kind == FT_K_UNION) {
                float m = fminf(f.acc, val);  // :33  (first child: acc = +inf)
                add_flops<COUNT>(acc, FLOPS_MINMAX);
                uint32_t i = f.idx + 1, child;
                if (union_scan<COUNT>(S, p, f.aux, i, m, child, acc)) {
                    val = m;
                    sp--;
                } else {
                    f.idx = i;
                    f.acc = m;
                    cur = child;
                    descend = true;
                }
            } else if (kind == FT_K_SUBTRACT) {
                const float4 h = ld4(S.nodes + off + 1);
                if (f.idx == 0u) {
                    f.idx = 1u;
                    f.acc = val;  // a.Distance position
                    cur = __float_as_uint(h.y);
                    descend = true;
                } else {
                    val = fmaxf(-val, f.acc);  // MathF.Max(-(b p), a p)
                    add_flops<COUNT>(acc, 2);
                    sp--;
                }
            } else if (kind == FT_K_INTERSECT) {
                const float4 h = ld4(S.nodes + off + 1);
                const uint32_t n = __float_as_uint(h.x), base = __float_as_uint(h.y);
                float mx = (f.idx == 0u) ? val : fmaxf(f.acc, val);  // :59 / :63
                uint32_t i = f.idx + 1;
                bool done = true;
                for (; i < n; i++) {
                    const uint32_t child = __ldg(S.links + base + i);
                    const float4 b0 = ld4(S.nodes + FT_WORD_OFF(child));
                    const float dc = distance(xyz(b0), p);
                    add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
                    if (mx < dc + b0.w) {  // :62 SdfBoundary.getMaxDistance
                        if (FT_WORD_KIND(child) == FT_K_SPHERE) {
                            mx = fmaxf(mx, dc - b0.w);
                            add_flops<COUNT>(acc, 2);
                        } else if (FT_IS_PRIMITIVE(child)) {
                            mx = fmaxf(mx, prim_distance<COUNT>(S, child, b0, p, acc));
                            add_flops<COUNT>(acc, FLOPS_MINMAX);
                        } else {
                            f.idx = i;
                            f.acc = mx;
                            cur = child;
                            done = false;
                            break;
                        }
                    }
                }
                if (done) {
                    val = mx;
                    sp--;
                } else {
                    descend = true;
                }
            } else {  // smooth
                const float4 h = ld4(S.nodes + off + 1);
                const uint32_t n = __float_as_uint(h.z), base = __float_as_uint(h.w);
                const float sum = f.acc + expf(h.y * val);  // :80
                add_flops<COUNT>(acc, FLOPS_SMOOTH_TERM);
                const uint32_t i = f.idx + 1;
                if (i < n) {
                    f.idx = i;
                    f.acc = sum;
                    cur = __ldg(S.links + base + i);
                    descend = true;
                } else {
                    val = -logf(sum) * h.x;  // :82
                    add_flops<COUNT>(acc, 3);
                    sp--;
                }
            }
        }
    }
}

// out-of-line copy for the rare callers (composite children inside the material arg-min)
template <bool COUNT>
__device__ __noinline__ float eval_form_call(const DevScene &S, uint32_t root, V3 p, Acc &acc) {
    return eval_form<COUNT>(S, root, p, acc);
}

// ---- material arg-min: the leaf SdfObject whose material colours p (SdfObject.fs:27-46, 50-64)
template <bool COUNT>
__device__ int material_leaf(const DevScene &S, V3 p, Acc &acc) {
    int o = S.root_object;
    for (;;) {
        const int4 rec = __ldg(S.objects + o);
        if (rec.x == FT_OBJECT_LEAF) return o;
        if (rec.x != FT_OBJECT_UNION) {  // subtract / intersect keep the base object's material
            o = rec.z;
            continue;
        }
        uint32_t begin;
        float dtc;
        grid_cell<COUNT>(S, FT_WORD_OFF((uint32_t)rec.y), p, begin, dtc, acc);
        uint32_t best = begin;
        const Entry e0 = ld_entry(S, begin);
        const uint32_t w0 = e0.word();
        float m = FT_IS_PRIMITIVE(w0) ? prim_distance<COUNT>(S, w0, e0.bound(), p, acc)
                                      : eval_form_call<COUNT>(S, w0, p, acc);  // :30
        // the reference's loop starts again at item 0 (:34); re-testing it cannot change min or material
        for (uint32_t i = begin + 1;; i++) {
            const Entry e = ld_entry(S, i);
            add_flops<COUNT>(acc, FLOPS_LB_TEST);
            if (!(m > e.lb() - dtc)) break;
            const float4 b0 = e.bound();
            const uint32_t w = e.word();
            float bd;
            add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
            if (bound_test(m, b0, p, bd)) {
                const float dd = (FT_WORD_KIND(w) == FT_K_SPHERE) ? bd
                                 : FT_IS_PRIMITIVE(w)              ? prim_distance<COUNT>(S, w, b0, p, acc)
                                                                   : eval_form_call<COUNT>(S, w, p, acc);
                if (dd < m) {  // :41-43
                    m = dd;
                    best = i;
                }
            }
        }
        o = __ldg(S.obj_links + rec.z + ld_entry(S, best).child());
    }
}

// ---- escape proof (device_scene.h "Escape box"): true when the half-line {o + t d, t >= 0} certainly stays
// farther than eps from the escape box, so that no future Distance along the ray can be < eps.  Conservative:
// any NaN or doubt answers false.  The reciprocal may be approximate: the box carries a 1e-3 relative slack.
__device__ __forceinline__ bool ray_escapes(const DevScene &S, V3 o, V3 d, float eps) {
    const float ov[3] = {o.x, o.y, o.z}, dv[3] = {d.x, d.y, d.z};
    float tmin = 0.0f, tmax = __int_as_float(0x7f800000);
    bool apart = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = S.esc_lo[k] - eps, hi = S.esc_hi[k] + eps;
        if (dv[k] == 0.0f) {
            apart = apart || ov[k] < lo || ov[k] > hi;  // parallel to the slab and outside it
        } else {
            const float inv = __fdividef(1.0f, dv[k]);
            const float t1 = (lo - ov[k]) * inv, t2 = (hi - ov[k]) * inv;
            tmin = fmaxf(tmin, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    return apart || tmax < tmin;
}
__device__ __forceinline__ bool outside_box(const float lo[3], const float hi[3], V3 p) {
    return p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x >= hi[0] || p.y >= hi[1] || p.z >= hi[2];
}

// ---- Image.getUniformPixelPos + Camera.uniformPixelToRay (Image.fs:17-23, Camera.fs:44-54)
__device__ __forceinline__ V3 pixel_direction(const ft_camera &cam, int x, int y, float max_size) {
    const float u = (float)x / max_size, v = (float)y / max_size;
    const V3 fwd = V3{cam.forward[0], cam.forward[1], cam.forward[2]};
    const V3 rs = V3{cam.right_scaled[0], cam.right_scaled[1], cam.right_scaled[2]};
    const V3 us = V3{cam.up_scaled[0], cam.up_scaled[1], cam.up_scaled[2]};
    return normalize((fwd + rs * (u - 0.5f)) + us * (v - 0.5f));
}

}  // namespace FT_NS
