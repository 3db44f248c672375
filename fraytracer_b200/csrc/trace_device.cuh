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
constexpr int FLOPS_CONE = 52;  // EXTENSION primitive (no reference)
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

#ifndef FT_FMA_PREFILTER
#define FT_FMA_PREFILTER 1  // contracted sum of squares in the certain-fail filter of the second pruning test (the exact value is recomputed when the outcome is open).  Measured on B200: nothing while the kernel spilled (2.395 vs 2.408 ms); -2.6 % on config 2 with the lean lane state
#endif
#ifndef FT_SCAN_PTR
#define FT_SCAN_PTR 1  // the candidate scan walks a pointer (immediate offsets, one 64-bit add per two records) instead of re-deriving the address from an index
#endif
#ifndef FT_SCAN_UNROLL
#define FT_SCAN_UNROLL 4  // candidates per round of the scan loop (2 or 4).  Measured on B200 with the lean lane state (no spills either way): 4 is 1-2 % faster on configs 1, 2 and 5
#endif
#ifndef FT_CHAIN_MAT
#define FT_CHAIN_MAT 1  // the chain evaluator notes its arg-min candidate, so the material lookup of a hit needs no second scan.
                       // Bit-exact (all parity tests).  Measured on B200: while the kernel spilled 170 B the note's register cost
                       // more than the scan it saves (config 2 -2.8 %); with the lean lane state it is config 2 +9 %, config 5 +5 %
#endif
#ifndef FT_WITH_CONE
#define FT_WITH_CONE 0
#endif
// sub_lookup, round 2 (each measured on B200, config 2):
#ifndef FT_SUB_BIASED
#define FT_SUB_BIASED 1  // the halo offset of the three fine-cell coordinates is folded into the header pointer on the host
#endif
#ifndef FT_SUB_IDX_WALK
#define FT_SUB_IDX_WALK 1  // the candidate list is walked by a 32-bit position (two instructions and a register less per candidate)
#endif
#ifndef FT_SUB_TIE_FALLBACK
#define FT_SUB_TIE_FALLBACK 1  // exact ties between candidates go to the sequential scan instead of being resolved by index
#endif
#ifndef FT_SUB_PRESCALED
#define FT_SUB_PRESCALED 1  // the ray machine hands the evaluator the fine-cell coordinate (see chain_coord)
#endif
#ifndef FT_SUB_EARLY_RETURN
#define FT_SUB_EARLY_RETURN 1  // a failed certificate leaves sub_lookup at once (no boolean carried in a register to the end)
#endif
#ifndef FT_SUB_DEBUG_IN_KERNEL
#define FT_SUB_DEBUG_IN_KERNEL 0  // 1: the kernel itself tests DevScene::sub_debug after every certified evaluation
#endif
#ifndef FT_INLINE_PRIM
#define FT_INLINE_PRIM 0  // 1: primitives met by the candidate scan are evaluated inside the scan loop
#endif

struct Acc {
    unsigned long long flops;
    unsigned long long steps;
    unsigned long long long_flops = 0ull;  // COUNT only: flops / evaluations of positions outside the largest grid's box
    unsigned long long long_evals = 0ull, evals = 0ull;
};

template <bool COUNT>
__device__ __forceinline__ void add_flops(Acc &a, int n) {
    if (COUNT) a.flops += (unsigned)n;
}

__device__ __forceinline__ float4 ld4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ V3 xyz(float4 v) { return V3{v.x, v.y, v.z}; }

// ---- EXTENSION primitive, no reference implementation: exact capped-cone distance.  Compiled only into
// the builds that define FT_WITH_CONE: inside the one evaluator call site even the untaken branch cost scenes WITHOUT cones
// 4 % (config 1 and 5, measured on B200 -- code size and register allocation of the hot loop), inlined or as a call.
#if FT_WITH_CONE
__device__ __forceinline__ float cone_distance(const float4 *rec, V3 p) {
    const float4 r1 = ld4(rec + 1), r2 = ld4(rec + 2), r3 = ld4(rec + 3);
    const V3 ba = xyz(r2), pa = p - xyz(r1);
    const float ra = r1.w, rb = r2.w, baba = r3.x, rba = r3.y, k = r3.z;
    const float papa = dot(pa, pa);
    const float paba = dot(pa, ba) / baba;
    const float x2 = papa - (paba * paba) * baba;
    const float x = sqrtf(x2 > 0.0f ? x2 : 0.0f);
    const float cx = x - (paba < 0.5f ? ra : rb);
    const float cax = cx > 0.0f ? cx : 0.0f;
    const float cay = fabsf(paba - 0.5f) - 0.5f;
    const float fr = (rba * (x - ra) + paba * baba) / k;
    const float f = fr < 0.0f ? 0.0f : (fr > 1.0f ? 1.0f : fr);
    const float cbx = (x - ra) - f * rba;
    const float cby = paba - f;
    const float da = cax * cax + (cay * cay) * baba, db = cbx * cbx + (cby * cby) * baba;
    const float dm = sqrtf(da < db ? da : db);
    return (cbx < 0.0f && cay < 0.0f) ? -dm : dm;
}
#endif

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
#if FT_WITH_CONE
    if (kind == FT_K_CONE) {  // EXTENSION primitive
        add_flops<COUNT>(acc, FLOPS_CONE);
        return cone_distance(rec, p);
    }
#endif
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
// MathI.clamp 0 m (MathF.floor_i x) (SdfBoundary.fs:276-282, Math.fs:23,57) in five instructions: floor(x) lies in [0, 2^31)
// exactly when x does, the float -> int conversion is then exact, and everything else -- negative, too large (cvttss2si
// gives INT_MIN), NaN (INT_MIN as well) -- clamps to 0.  Same values as clamp_int(0, m, floor_to_int(x)) for every float.
__device__ __forceinline__ int cell_index(float x, int m) {
    const int i = __float2int_rd(x);
    return (x >= 0.0f && x < 2147483648.0f) ? (i < m ? i : m) : 0;
}
__device__ __forceinline__ void grid_cell_from(const DevScene &S, float4 g1, float4 g2, float4 g3, V3 p, uint32_t &begin, float &dtc) {
    const V3 c = (p - xyz(g1)) * xyz(g2);
    const int mx = __float_as_int(g3.x), my = __float_as_int(g3.y), mz = __float_as_int(g3.z);
    const int ix = cell_index(c.x, mx), iy = cell_index(c.y, my), iz = cell_index(c.z, mz);
    const uint32_t ci = __float_as_uint(g1.w) + (uint32_t)(ix * __float_as_int(g3.w) + iy * (mz + 1) + iz);
    const float4 cell = ld4(S.cells + ci);
    dtc = distance(p, xyz(cell));  // SdfForm.fs:25
    begin = __float_as_uint(cell.w);
}
template <bool COUNT>
__device__ __forceinline__ void grid_cell(const DevScene &S, uint32_t uoff, V3 p, uint32_t &begin, float &dtc, Acc &acc) {
    add_flops<COUNT>(acc, FLOPS_CELL_LOOKUP);
    grid_cell_from(S, ld4(S.nodes + uoff + 1), ld4(S.nodes + uoff + 2), ld4(S.nodes + uoff + 3), p, begin, dtc);
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
// s = |C - p|^2 is formed exactly like the reference; the root is only taken when the outcome is open.
// With u = max(m + R, 1e-15):  s > u^2 (1 + 2^-12)  implies  q = fl(sqrt(s)) > m + R  (u^2 >= 1e-30 is a normal
// float, so the 2^-13 margin on the root dwarfs the roundings of u, u^2, the constant product and q; when
// m + R < 1e-15, q > 1e-15 > m + R directly), hence q - R > m as reals and, rounding being monotone,
// fl(q - R) >= m: the reference's test is false.  NaN in m or R lands here too (fmaxf drops it) and the
// reference's comparison with NaN is false as well; NaN in s fails `s > ..` and takes the exact path.
// Otherwise bd = fl(fl(sqrt(s)) - R) is computed and compared as the reference does.  `bd` is only valid when
// the function returns true.
__device__ __forceinline__ bool bound_test(float m, float4 b0, V3 p, float &bd) {
    const V3 dl = xyz(b0) - p;
    const float u = fmaxf(m + b0.w, 1e-15f);
#if FT_FMA_PREFILTER
    // The certain-fail filter does not need the reference's rounding of s: three non-negative terms, so the
    // contracted sum and the reference's individually rounded one are both within 2^-21 (relative) of the real
    // value and of each other.  s' > u^2 (1 + 2^-12) gives s > u^2 (1 + 2^-13), which is the premise the proof
    // above needs (its margin was 2^-13 on the root already with the constant halved).  Explicit fma intrinsics:
    // this translation unit is compiled with -fmad=false, nothing else contracts.
    const float sf = __fmaf_rn(dl.z, dl.z, __fmaf_rn(dl.y, dl.y, dl.x * dl.x));
    if (sf > (u * u) * 1.000244140625f) return false;
    const float s = dot(dl, dl);
#else
    const float s = dot(dl, dl);
    if (s > (u * u) * 1.000244140625f) return false;
#endif
    bd = sqrtf(s) - b0.w;
    return m > bd;
}

// ---- union candidate scan (SdfForm.fs:27-33) from entry i with running minimum m.
// Returns true when the list is exhausted (early exit on the sorted LowerBound, see device_scene.h);
// false when entry i is a child whose Distance must be evaluated by the caller (anything but a sphere:
// a sphere's Distance is the bound test's own value).  Started at the cell's first entry with m = +inf it
// also performs the reference's unconditional first evaluation (SdfForm.fs:26): +inf passes both tests.
// Two records are in flight (ping-pong, no register rotation); the list is closed by a sentinel and the
// array by two padding records, so the speculative loads stay in bounds.
template <bool COUNT>
__device__ __forceinline__ bool union_scan(const DevScene &S, V3 p, float dtc, uint32_t &i, float &m, uint32_t &child, Acc &acc) {
#define FT_SCAN_ONE(E)                                                                    \
    {                                                                                     \
        add_flops<COUNT>(acc, FLOPS_LB_TEST);                                             \
        if (!(m > (E).lb() - dtc)) return true;                                           \
        float bd;                                                                         \
        add_flops<COUNT>(acc, FLOPS_BOUND_TEST);                                          \
        if (bound_test(m, (E).bound(), p, bd)) { /* m > getMinDistance boundary p */      \
            if (FT_WORD_KIND((E).word()) == FT_K_SPHERE) {                                \
                m = bd; /* Min(m, Distance) with Distance == bd < m */                    \
                add_flops<COUNT>(acc, FLOPS_MINMAX);                                      \
            } else if (FT_INLINE_PRIM && FT_IS_PRIMITIVE((E).word())) {                   \
                m = fminf(m, prim_distance<COUNT>(S, (E).word(), (E).bound(), p, acc));   \
                add_flops<COUNT>(acc, FLOPS_MINMAX);                                      \
            } else {                                                                      \
                child = (E).word();                                                       \
                return false;                                                             \
            }                                                                             \
        }                                                                                 \
    }
    Entry e0 = ld_entry(S, i);
    for (;;) {
        const Entry e1 = ld_entry(S, i + 1);
        FT_SCAN_ONE(e0)
        i++;
        e0 = ld_entry(S, i + 1);
        FT_SCAN_ONE(e1)
        i++;
    }
#undef FT_SCAN_ONE
}

struct Frame {
    uint32_t word;
    uint32_t idx;
    float acc;
    float aux;
};

// ---- Form.Distance position for an arbitrary compiled form (the closure of Types.fs:42).
// A resumable evaluate/feed loop: "evaluate" turns `cur` into a value (the ONE primitive site) or opens a
// composite frame; "feed" hands the value to the innermost open composite, which either asks for its next
// child or closes.  Lanes of a warp therefore meet at the primitive site for their k-th child evaluation
// instead of evaluating primitives at different iterations of their private scan loops.  The innermost
// frame lives in registers; only deeper nesting touches the local-memory stack.
template <bool COUNT>
__device__ __forceinline__ float eval_form(const DevScene &S, uint32_t root, V3 p, Acc &acc) {
    Frame stack[FT_MAX_DEPTH];
    Frame top{0u, 0u, 0.f, 0.f};
    int depth = 0;  // open composite frames; frames 0..depth-2 are in stack[], frame depth-1 is `top`
    uint32_t cur = root;
    float val = 0.f;
    bool have_val = false;
    for (;;) {
        if (!have_val) {
            const uint32_t kind = FT_WORD_KIND(cur), off = FT_WORD_OFF(cur);
            if (kind < 8u) {
                val = prim_distance<COUNT>(S, cur, ld4(S.nodes + off), p, acc);
                have_val = true;
            } else {
                if (depth > 0) stack[depth - 1] = top;
                depth++;
                const float4 h = ld4(S.nodes + off + 1);
                if (kind == FT_K_UNION) {  // SdfForm.fs:22-34
                    uint32_t begin;
                    float dtc;
                    grid_cell<COUNT>(S, off, p, begin, dtc, acc);
                    top = Frame{cur, begin - 1u, __int_as_float(0x7f800000), dtc};
                    val = __int_as_float(0x7f800000);  // feed(+inf): the scan starts at the first entry
                    have_val = true;
                } else if (kind == FT_K_SUBTRACT) {  // SdfForm.fs:45-47
                    top = Frame{cur, 0u, 0.f, 0.f};
                    cur = __float_as_uint(h.x);
                } else if (kind == FT_K_INTERSECT) {  // SdfForm.fs:58-64
                    top = Frame{cur, 0u, 0.f, 0.f};
                    cur = __ldg(S.links + __float_as_uint(h.y));
                } else {  // FT_K_SMOOTH  SdfForm.fs:75-82
                    top = Frame{cur, 0u, 0.f, 0.f};
                    cur = __ldg(S.links + __float_as_uint(h.w));
                }
            }
        } else {
            if (depth == 0) return val;
            const uint32_t kind = FT_WORD_KIND(top.word), off = FT_WORD_OFF(top.word);
            bool closed = false;
            if (kind == FT_K_UNION) {
                float m = fminf(top.acc, val);  // SdfForm.fs:33
                add_flops<COUNT>(acc, FLOPS_MINMAX);
                uint32_t i = top.idx + 1u, child;
                if (union_scan<COUNT>(S, p, top.aux, i, m, child, acc)) {
                    val = m;
                    closed = true;
                } else {
                    top.idx = i;
                    top.acc = m;
                    cur = child;
                    have_val = false;
                }
            } else if (kind == FT_K_SUBTRACT) {
                const float4 h = ld4(S.nodes + off + 1);
                if (top.idx == 0u) {
                    top.idx = 1u;
                    top.acc = val;  // a.Distance position
                    cur = __float_as_uint(h.y);
                    have_val = false;
                } else {
                    val = fmaxf(-val, top.acc);  // MathF.Max(-(b p), a p)
                    add_flops<COUNT>(acc, 2);
                    closed = true;
                }
            } else if (kind == FT_K_INTERSECT) {
                const float4 h = ld4(S.nodes + off + 1);
                const uint32_t n = __float_as_uint(h.x), base = __float_as_uint(h.y);
                float mx = (top.idx == 0u) ? val : fmaxf(top.acc, val);  // :59 / :63
                uint32_t i = top.idx + 1u;
                closed = true;
                for (; i < n; i++) {
                    const uint32_t child = __ldg(S.links + base + i);
                    const float4 b0 = ld4(S.nodes + FT_WORD_OFF(child));
                    const float dc = distance(xyz(b0), p);
                    add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
                    if (mx < dc + b0.w) {  // :62 SdfBoundary.getMaxDistance
                        if (FT_WORD_KIND(child) == FT_K_SPHERE) {
                            mx = fmaxf(mx, dc - b0.w);
                            add_flops<COUNT>(acc, 2);
                        } else {
                            top.idx = i;
                            top.acc = mx;
                            cur = child;
                            have_val = false;
                            closed = false;
                            break;
                        }
                    }
                }
                if (closed) val = mx;
            } else {  // smooth
                const float4 h = ld4(S.nodes + off + 1);
                const uint32_t n = __float_as_uint(h.z), base = __float_as_uint(h.w);
                const float sum = top.acc + expf(h.y * val);  // :80
                add_flops<COUNT>(acc, FLOPS_SMOOTH_TERM);
                const uint32_t i = top.idx + 1u;
                if (i < n) {
                    top.idx = i;
                    top.acc = sum;
                    cur = __ldg(S.links + base + i);
                    have_val = false;
                } else {
                    val = -logf(sum) * h.x;  // :82
                    add_flops<COUNT>(acc, 3);
                    closed = true;
                }
            }
            if (closed) {
                depth--;
                if (depth > 0) top = stack[depth - 1];
            }
        }
    }
}

// ---- specialised evaluator for chain-shaped root forms (device_scene.h "Chain program"):
//   v = U(p)  (SdfForm.fs:22-34),  then  v = op_k(v)  for the intersect / subtract operations above U.
// Bit-identical to eval_form on such scenes; it only drops the interpreter's dispatch.  The candidate loop is
// "scan until a child must be evaluated, converge, evaluate": lanes share the primitive-evaluation code.
// `best_lo` receives the low 32 address bits of the candidate record that set the running minimum last (the arg-min
// of SdfObject.fs:27-46, see DevScene::chain_mat_valid); callers that do not need it pass a dummy and the stores vanish.
// ---- U(p) through the nearest-candidate sub-grid (device_scene.h, subgrid_build.cu; DESIGN.md "sub-grid").
//
// The reference's value (SdfForm.fs:22-34) is  m_seq = min over E_ref of D_i(p),  E_ref the candidates its sequential scan
// evaluates: the scan ends at the first K with m <= fl(lb_K - dtc) (sorted list, m never grows) and skips i < K with
// m <= bd_i = fl(|C_i - p| - R_i).  Its arg-min (the material leaf, SdfObject.fs:27-46) is the FIRST index attaining m_seq.
//
// Claim: let W be any subset of the cell's list that contains that arg-min i*, M = min over W of D_i(p) and a the smallest
// index in W attaining M.  If  (1) M >= fl(lb_a - dtc)  and  (2) M >= bd_a  (trivial for a sphere: D = bd),  then M == m_seq
// bit for bit and a == i*.   Proof: i* in W gives M <= m_seq.  If a < K: either the scan evaluates a (then D_a >= m_seq), or it
// skips a because m_{a-1} <= bd_a <= D_a (by (2)), and m_{a-1} >= m_seq.  If a >= K: D_a >= fl(lb_a - dtc) (by (1)) >=
// fl(lb_K - dtc) >= m_{K-1} = m_seq.  So M == m_seq, and a == i* because an index below i* attaining m_seq would either be
// evaluated by the scan (contradicting "first") or be skipped with m_{a-1} <= D_a = m_seq, i.e. the minimum was already
// attained even earlier.  No order of evaluation is involved: W may be visited nearest-first.
//
// W is the sub-cell's list: subgrid_build.cu drops candidate i only when D_i > D_j + margin for some j of the same cell list
// everywhere in the sub-cell, and m_seq(p) <= D_j(p) + (a few ulps) for EVERY j of the list (evaluated: the scan takes the
// minimum; skipped: m <= bd_j <= D_j; beyond K: m_seq <= fl(lb_j - dtc) <= D_j up to rounding), so a dropped candidate cannot
// be i*.  Certificates (1) and (2) are checked on the final arg-min only, first by a square-root-free sufficient test, then
// exactly; if one fails -- only possible within rounding of a tie AND of a degenerate configuration -- the scan itself runs.
//
// certain_fail_up: true only if  M > bd  is certainly false, i.e. D_i >= bd_i - ulps > M: the candidate cannot lower M (nor tie).
// `rup` >= R (1 + 2^-12) is stored in the record.  u = fl(M (1 + 2^-12) + rup) >= (M + R)(1 + 2^-12)(1 - 2^-24) whenever
// M + R >= 0, and sf > fl(u u + 1e-30) then gives sqrt(s) > (M + R)(1 + 2^-13): q = fl(sqrt(fl(s))) > M + R, so fl(q - R) >= M.
// (1e-30 keeps the comparison meaningful when u u underflows; M + R < 0 makes the reference's test false for every s;
// M = +inf gives u = +inf and the filter never fires.)  sf is the contracted sum of squares, within 2^-21 of the reference's.
__device__ __forceinline__ bool certain_fail_up(float M, float4 b0, float rup, V3 p) {
    const V3 dl = xyz(b0) - p;
    const float sf = __fmaf_rn(dl.z, dl.z, __fmaf_rn(dl.y, dl.y, dl.x * dl.x));
    const float u = __fmaf_rn(M, 1.000244140625f, rup);
    return sf > __fmaf_rn(u, u, 1e-30f);
}

// The cell coordinate of p in U's spatial lookup, c = (p - aabbMin) * cellSizeInv (SdfBoundary.fs:276, header from the constant
// bank), and whether it lies in the "fast region" -- the sub-grid's covered region clipped to the escape box (DevScene::fast_lo).
// One box test serves two decisions there: the evaluator goes straight to the sub-grid, and the ray machine skips the escape
// pre-check (a position inside the escape box cannot pass it).
// FT_SUB_PRESCALED: the coordinate handed from the ray machine to the evaluator is the FINE-cell coordinate
// f = (p - aabbMin) * (cellSizeInv * sub_scale) = c * sub_scale bit for bit (sub_scale is a power of two: scaling commutes with
// the rounding of the product, underflow aside -- and a product that underflows lies within 1e-37 of the box's min face, where
// either neighbouring fine cell draws from the same reference cell with a margin of millimetres), so the fast path needs
// neither c nor the multiplication by sub_scale; the fast region is kept in the same units.  The cold paths form c themselves.
__device__ __forceinline__ V3 cell_coord(const DevScene &S, V3 p) { return (p - xyz(S.chain_g1)) * xyz(S.chain_g2); }
#if FT_SUB_PRESCALED
__device__ __forceinline__ V3 chain_coord(const DevScene &S, V3 p) { return (p - xyz(S.chain_g1)) * xyz(S.chain_g2s); }
__device__ __forceinline__ bool in_fast_region(const DevScene &S, V3 f) {
    return f.x >= S.fast_lo_s[0] && f.y >= S.fast_lo_s[1] && f.z >= S.fast_lo_s[2] && f.x < S.fast_hi_s[0] &&
           f.y < S.fast_hi_s[1] && f.z < S.fast_hi_s[2];
}
#else
__device__ __forceinline__ V3 chain_coord(const DevScene &S, V3 p) { return cell_coord(S, p); }
__device__ __forceinline__ bool in_fast_region(const DevScene &S, V3 c) {
    return c.x >= S.fast_lo[0] && c.y >= S.fast_lo[1] && c.z >= S.fast_lo[2] && c.x < S.fast_hi[0] && c.y < S.fast_hi[1] &&
           c.z < S.fast_hi[2];
}
#endif

// `g` = the fine-cell coordinates of p (see DevScene): the caller has checked that p lies in the covered region
template <bool COUNT>
__device__ __forceinline__ bool sub_lookup(const DevScene &S, V3 p, int gx, int gy, int gz, float &m_out, uint32_t &best_out,
                                           Acc &acc) {
    // header: {(first << 8) | count, the nearest candidate}: the first record's load depends on one load only
#if FT_SUB_BIASED
    // `g` without the halo offset: the header pointer is biased by it on the host (DevScene::sub_hdr_biased)
    const uint2 hdr = __ldg(S.sub_hdr_biased + ((gx * S.sub_ny + gy) * S.sub_nz + gz));
#else
    const uint2 hdr = __ldg(S.sub_hdr + (uint32_t)((gx * S.sub_ny + gy) * S.sub_nz + gz));
#endif
    const uint32_t n = hdr.x & 255u;
    if (n == 0u) return false;  // no list for this sub-cell (too long to encode): the scan runs
    add_flops<COUNT>(acc, 9);
    float M = __int_as_float(0x7f800000);
    uint32_t best = 0xffffffffu;
    uint32_t idx = hdr.y & 0xffffffu;
#if FT_SUB_IDX_WALK
    // candidates 1 .. n-1 by position in sub_idx (a 32-bit cursor and its end instead of a 64-bit pointer and a counter)
    uint32_t pos = hdr.x >> 8;
    const uint32_t end = pos + n - 1u;
    for (;;) {
#else
    const uint32_t *lp = S.sub_idx + (hdr.x >> 8);  // candidates 1 .. n-1
    for (uint32_t k = 0;;) {
#endif
        const Entry e = ld_entry(S, idx);
#if !FT_SUB_IDX_WALK
        const bool more = ++k < n;
#endif
        add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
        if (!certain_fail_up(M, e.bound(), e.b.w, p)) {
            float d;
            if (FT_WORD_KIND(e.word()) == FT_K_SPHERE) {
                d = distance(xyz(e.bound()), p) - e.b.y;  // SdfForm.fs:127-129, the reference's operation order
            } else {
                d = prim_distance<COUNT>(S, e.word(), e.bound(), p, acc);
            }
            add_flops<COUNT>(acc, FLOPS_MINMAX);
#if FT_SUB_TIE_FALLBACK
            // an exact tie at the running minimum (the arg-min would be the smaller index) invalidates `best` until a strictly
            // smaller Distance turns up: a tie that survives leaves the evaluation to the scan
            if (d <= M) {
                best = d < M ? idx : 0xffffffffu;
                M = d;
            }
#else
            if (d < M || (d == M && idx < best)) {
                M = d;
                best = idx;
            }
#endif
        }
#if FT_SUB_IDX_WALK
        if (pos == end) break;
        // (the address is formed in inline PTX: the compiler otherwise turns the cursor back into a 64-bit pointer and a counter)
        asm("{\n\t"
            ".reg .u64 a;\n\t"
            "mad.wide.u32 a, %1, 4, %2;\n\t"
            "ld.global.nc.u32 %0, [a];\n\t"
            "}"
            : "=r"(idx)
            : "r"(pos), "l"(S.sub_idx));
        pos++;
#else
        if (!more) break;
        idx = __ldg(lp + k - 1u);  // (prefetching it one candidate ahead costs a register: measured 0.792 -> 0.821 ms)
#endif
    }
    if (best == 0xffffffffu) return false;  // nothing comparable (NaN positions): the scan decides
    // certificates on the arg-min
    const Entry eb = ld_entry(S, best);
    add_flops<COUNT>(acc, 12);
    const float x = eb.lb() - M;
#if !FT_SUB_EARLY_RETURN
    bool ok = true;
#endif
    // (1)  M >= fl(lb - dtc).  Certainly true when x = fl(lb - M) <= dq, dq = the header's lower bound of dtc over the whole fine
    // cell (rounded down to its 8-bit quantum and lowered by one more, which dwarfs the roundings of x, dtc and the
    // difference): lb - dtc <= lb - dq <= M.  The comparison runs on 1 + x / (256 quanta) against the float whose top mantissa
    // bits are q (no integer-to-float conversion).  Otherwise: certainly true when |p - centre|^2 > x^2 (1 + 2^-10); else exact.
    if (__fmaf_rn(x, S.sub_dtc_inv, 1.0f) > __uint_as_float(0x3f800000u | ((hdr.y >> 24) << 15))) {
        // the reference's cell, clamp(floor(c)) per axis (SdfBoundary.fs:276-282), from the position again rather than from the
        // fine-cell coordinates: those need not stay live across the candidate loop (three registers where the kernel spills)
        const V3 c = cell_coord(S, p);
        const int mx = __float_as_int(S.chain_g3.x), my = __float_as_int(S.chain_g3.y), mz = __float_as_int(S.chain_g3.z);
        const int ix = cell_index(c.x, mx), iy = cell_index(c.y, my), iz = cell_index(c.z, mz);
        const float4 cell = ld4(S.cells + (__float_as_uint(S.chain_g1.w) +
                                           (uint32_t)(ix * __float_as_int(S.chain_g3.w) + iy * (mz + 1) + iz)));
        const V3 dc = p - xyz(cell);
        const float dtc2 = __fmaf_rn(dc.z, dc.z, __fmaf_rn(dc.y, dc.y, dc.x * dc.x));
#if FT_SUB_EARLY_RETURN
        if (!(dtc2 > (x * x) * 1.0009765625f) && !(M >= eb.lb() - distance(p, xyz(cell)))) return false;
#else
        if (!(dtc2 > (x * x) * 1.0009765625f)) ok = M >= eb.lb() - distance(p, xyz(cell));
#endif
    }
    if (FT_WORD_KIND(eb.word()) != FT_K_SPHERE) {  // (2)  M >= bd: certainly true when s < (M + R)^2 (1 - 2^-16), M + R > 0
        const float4 b0 = eb.bound();
        const V3 dl = xyz(b0) - p;
        const float sf = __fmaf_rn(dl.z, dl.z, __fmaf_rn(dl.y, dl.y, dl.x * dl.x));
        const float y = M + b0.w;
#if FT_SUB_EARLY_RETURN
        if (!(y > 0.0f && sf < (y * y) * 0.9999847412109375f) && !(M >= sqrtf(dot(dl, dl)) - b0.w)) return false;
#else
        if (!(y > 0.0f && sf < (y * y) * 0.9999847412109375f)) ok = ok && (M >= sqrtf(dot(dl, dl)) - b0.w);
#endif
    }
    // ("distrust every certificate", sub_debug bit 0, is applied by the host: it detaches the sub-grid)
#if FT_SUB_EARLY_RETURN
#if FT_SUB_DEBUG_IN_KERNEL
    if (S.sub_debug & 1) return false;
#endif
#elif FT_SUB_DEBUG_IN_KERNEL
    if (!ok || (S.sub_debug & 1)) return false;
#else
    if (!ok) return false;
#endif
    m_out = M;
    best_out = best;
    return true;
}

// `cf` = chain_coord(S, p), `fast` = in_fast_region(S, cf)
template <bool COUNT>
__device__ __forceinline__ float eval_chain_c(const DevScene &S, V3 p, V3 cf, bool fast, Acc &acc, uint32_t &best_lo) {
    uint32_t i;
    float dtc;
    add_flops<COUNT>(acc, FLOPS_CELL_LOOKUP);
    float m = __int_as_float(0x7f800000);  // +inf passes both tests: the first entry is evaluated unconditionally (:26)
    bool certified = false;
    {
        // SdfBoundary.fs:276-282 with U's header from the constant bank (see grid_cell_from)
        const float4 g1 = S.chain_g1, g3 = S.chain_g3;
#if FT_SUB_PRESCALED
        bool covered = fast;
        if (!fast) {
            const V3 c = cell_coord(S, p);  // the reference's cell coordinate: formed only off the fast path
            covered = S.sub_valid && c.x >= S.sub_lo[0] && c.y >= S.sub_lo[1] && c.z >= S.sub_lo[2] && c.x < S.sub_hi[0] &&
                      c.y < S.sub_hi[1] && c.z < S.sub_hi[2];
        }
        if (covered) {
            uint32_t best = 0u;
            float ms = 0.f;
            static_assert(FT_SUB_BIASED, "FT_SUB_PRESCALED needs the biased header pointer");
            certified = sub_lookup<COUNT>(S, p, __float2int_rd(cf.x), __float2int_rd(cf.y), __float2int_rd(cf.z), ms, best, acc);
#else
        const V3 c = cf;
        if (fast || (S.sub_valid && c.x >= S.sub_lo[0] && c.y >= S.sub_lo[1] && c.z >= S.sub_lo[2] && c.x < S.sub_hi[0] &&
                     c.y < S.sub_hi[1] && c.z < S.sub_hi[2])) {
            uint32_t best = 0u;
            float ms = 0.f;
#if FT_SUB_BIASED
            certified = sub_lookup<COUNT>(S, p, __float2int_rd(c.x * S.sub_scale), __float2int_rd(c.y * S.sub_scale),
                                          __float2int_rd(c.z * S.sub_scale), ms, best, acc);
#else
            certified = sub_lookup<COUNT>(S, p, __float2int_rd(c.x * S.sub_scale) + S.sub_off,
                                          __float2int_rd(c.y * S.sub_scale) + S.sub_off,
                                          __float2int_rd(c.z * S.sub_scale) + S.sub_off, ms, best, acc);
#endif
#endif
            if (certified) {
                m = ms;
#if FT_CHAIN_MAT
                best_lo = (uint32_t)(uintptr_t)(S.entries + 2ull * best);
#endif
            }
        }
        if (!certified) {
#if FT_SUB_PRESCALED
            const V3 c = cell_coord(S, p);
#endif
            const int mx = __float_as_int(g3.x), my = __float_as_int(g3.y), mz = __float_as_int(g3.z);
            const int ix = cell_index(c.x, mx), iy = cell_index(c.y, my), iz = cell_index(c.z, mz);
            const uint32_t ci = __float_as_uint(g1.w) + (uint32_t)(ix * __float_as_int(g3.w) + iy * (mz + 1) + iz);
            const float4 cell = ld4(S.cells + ci);
            dtc = distance(p, xyz(cell));  // SdfForm.fs:25
            i = __float_as_uint(cell.w);
        }
    }
    if (!certified) {
#if FT_SCAN_PTR
    // `ep` points at the record the unrolled loop started its current round with; the record under test is ep + 2 K
    // (K = 0, 1: an immediate offset).  The only pointer arithmetic inside the loop is one 64-bit add per two records;
    // an exit notes K in `kx` and `ep` is brought up to date outside.
    const float4 *ep = S.entries + 2ull * i;
#define FT_LD_ENTRY(K) Entry{__ldg(ep + 2 * (K)), __ldg(ep + 2 * (K) + 1)}
#if FT_CHAIN_MAT
#define FT_NOTE_BEST(K) best_lo = (uint32_t)(uintptr_t)(ep + 2 * (K))
#else
#define FT_NOTE_BEST(K) (void)0
#endif
#else
#define FT_LD_ENTRY(K) ld_entry(S, i + (K))
#define FT_NOTE_BEST(K) best_lo = (uint32_t)(uintptr_t)(S.entries + 2ull * (i + (K)))
#endif
    for (;;) {
        uint32_t w = 0u;
        float4 b0{0.f, 0.f, 0.f, 0.f};
        bool need = false;
        int kx = 0;
        // scan (SdfForm.fs:27-33); spheres are resolved in place (their Distance is the bound test's value);
        // two records in flight, ping-pong, so that the prefetch costs no register rotation
#define FT_CHAIN_ONE(E, K)                                                    \
    {                                                                         \
        add_flops<COUNT>(acc, FLOPS_LB_TEST);                                 \
        if (!(m > (E).lb() - dtc)) {                                          \
            kx = (K);                                                         \
            break;                                                            \
        }                                                                     \
        float bd;                                                             \
        add_flops<COUNT>(acc, FLOPS_BOUND_TEST);                              \
        if (bound_test(m, (E).bound(), p, bd)) {                              \
            if (FT_WORD_KIND((E).word()) == FT_K_SPHERE) {                    \
                m = bd;                                                       \
                FT_NOTE_BEST(K);                                              \
                add_flops<COUNT>(acc, FLOPS_MINMAX);                          \
            } else {                                                          \
                need = true;                                                  \
                w = (E).word();                                               \
                b0 = (E).bound();                                             \
                kx = (K);                                                     \
                break;                                                        \
            }                                                                 \
        }                                                                     \
    }
        Entry e0 = FT_LD_ENTRY(0);
#if FT_SCAN_UNROLL == 4
        Entry e1 = FT_LD_ENTRY(1);
        for (;;) {
            FT_CHAIN_ONE(e0, 0)
            e0 = FT_LD_ENTRY(2);
            FT_CHAIN_ONE(e1, 1)
            e1 = FT_LD_ENTRY(3);
            FT_CHAIN_ONE(e0, 2)
            e0 = FT_LD_ENTRY(4);
            FT_CHAIN_ONE(e1, 3)
            e1 = FT_LD_ENTRY(5);
#if FT_SCAN_PTR
            ep += 8;
#else
            i += 4;
#endif
        }
#else
        for (;;) {
            const Entry e1 = FT_LD_ENTRY(1);
            FT_CHAIN_ONE(e0, 0)
            e0 = FT_LD_ENTRY(2);
            FT_CHAIN_ONE(e1, 1)
#if FT_SCAN_PTR
            ep += 4;
#else
            i += 2;
#endif
        }
#endif
#undef FT_CHAIN_ONE
        if (!need) break;
        {
            const float dprim = prim_distance<COUNT>(S, w, b0, p, acc);
            if (dprim < m) {  // == Min(m, dprim) (SdfForm.fs:33); `d < min` is also the arg-min's update rule (SdfObject.fs:41)
                m = dprim;
                FT_NOTE_BEST(kx);
            }
        }
        add_flops<COUNT>(acc, FLOPS_MINMAX);
#if FT_SCAN_PTR
        ep += 2 * (kx + 1);
#else
        i += (uint32_t)kx + 1u;
#endif
    }
#undef FT_LD_ENTRY
#undef FT_NOTE_BEST
    }  // !certified
    float v = m;
    // (measured: also serving the chain operations' operands from the constant bank, unrolled, made config 1 4 % slower
    // -- four inlined copies of the primitive evaluator)
    for (int k = 0; k < S.chain_n_ops; k++) {
        const uint32_t op = __ldg(S.links + S.chain_ops + 2 * k), w = __ldg(S.links + S.chain_ops + 2 * k + 1);
        const float4 b0 = ld4(S.nodes + FT_WORD_OFF(w));
        if (op == FT_OP_INTERSECT) {
            const float dc = distance(xyz(b0), p);
            add_flops<COUNT>(acc, FLOPS_BOUND_TEST);
            if (v < dc + b0.w) {  // SdfForm.fs:62
                const float dd = (FT_WORD_KIND(w) == FT_K_SPHERE) ? dc - b0.w : prim_distance<COUNT>(S, w, b0, p, acc);
                v = fmaxf(v, dd);
                add_flops<COUNT>(acc, 2);
            }
        } else {
            const float dd = (FT_WORD_KIND(w) == FT_K_SPHERE) ? distance(xyz(b0), p) - b0.w
                                                              : prim_distance<COUNT>(S, w, b0, p, acc);
            v = fmaxf(-dd, v);  // SdfForm.fs:45-47
            add_flops<COUNT>(acc, 2);
        }
    }
    return v;
}

template <bool COUNT>
__device__ __forceinline__ float eval_chain(const DevScene &S, V3 p, Acc &acc, uint32_t &best_lo) {
    const V3 c = chain_coord(S, p);
    return eval_chain_c<COUNT>(S, p, c, in_fast_region(S, c), acc, best_lo);
}

// out-of-line copy of the chain evaluator for callers outside the hot loop (the once-per-warp probe of the camera position in
// the extension kernels).  (Measured and rejected on B200: out-of-line copies for the cold call sites INSIDE the loop -- the
// fallback scan, the chain operations' operands, the material arg-min -- to shrink the instruction footprint: no gain on the
// hot path (config 2 0.831 = 0.831 ms, config 1 0.669 vs 0.641 ms), and the extension kernels lost 8-30 %: live registers
// are spilled around every call.)
template <bool COUNT>
__device__ __noinline__ float eval_chain_call(const DevScene &S, V3 p, Acc &acc) {
    uint32_t best = 0u;
    return eval_chain<COUNT>(S, p, acc, best);
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
        float m = eval_form_call<COUNT>(S, w0, p, acc);  // :30
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
                const float dd = (FT_WORD_KIND(w) == FT_K_SPHERE) ? bd : eval_form_call<COUNT>(S, w, p, acc);
                if (dd < m) {  // :41-43
                    m = dd;
                    best = i;
                }
            }
        }
        o = __ldg(S.obj_links + rec.z + ld_entry(S, best).child());
    }
}

// ---- the leaf object from the hit evaluation's arg-min candidate (DevScene::chain_mat_valid): `best_lo` are the low address
// bits of that candidate's record, p the hit point (its cell is looked up again to anchor the 32-bit difference)
__device__ __forceinline__ int chain_leaf_from_best(const DevScene &S, V3 p, uint32_t best_lo) {
    uint32_t begin;
    float dtc;
    grid_cell_from(S, S.chain_g1, S.chain_g2, S.chain_g3, p, begin, dtc);
    const uint32_t first_lo = (uint32_t)(uintptr_t)(S.entries + 2ull * begin);
    const uint32_t idx = begin + ((best_lo - first_lo) >> 5);  // 32-byte records; a cell's list is far shorter than 2^27
    return __ldg(S.obj_links + S.chain_mat_links + ld_entry(S, idx).child());
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
