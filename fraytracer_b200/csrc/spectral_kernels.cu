// spectral_kernels.cu -- EXTENSION (BASELINE config 3): spectral / refractive ray-test as a wavefront pipeline.
//
// NOT reference behaviour.  The reference has no wavelengths, no glass and no rays spawned by materials
// (SURVEY.md section 0).  The spec is this repository's own, anchored in the reference's dead
// code: Light.fresnel (src/FrayTracer/Light.fs:30-59, used as written) and the frequency bands of Light.Colors
// (Light.fs:12-26).  Everything the reference does define is reused unchanged: the march loop SdfForm.tryTrace
// (SdfForm.fs:93-104), normalFromRay (SdfForm.fs:106-115), the material arg-min (SdfObject.fs:27-46), the light
// loop and Lambert sum (SdfScene.fs:12-28, SdfLight.fs:6-42).  Parity target: oracle/ft_oracle.c
// fto_render_spectral, bit for bit (this file is compiled with -fmad=false like the strict hot path).
//
// Pipeline (one launch each, no host synchronisation except one 4-byte read after wave 0):
//   wave 0   spectral_trace_kernel<PRIMARY>  one primary ray per pixel, shared by all wavelengths.  Misses and hits
//            on solid materials are resolved for every wavelength at once and written to the image; hits on glass
//            are appended to the hit queue (warp ballot + popc ranks, one atomicAdd per warp).
//   stage M  spectral_spawn_kernel           the material stage: one thread per (glass hit, wavelength) evaluates
//            the dispersive index, Light.fresnel and the Philox draw, and writes the continuation path.
//   wave w   spectral_trace_kernel<PATH>     one thread per live path: signed march, then resolve (miss / solid),
//            or interact again and append to the next wave's queue (same warp compaction).
//   resolve  spectral_resolve_kernel         per glass pixel: fixed-order sum of its wavelength slots -> RGB.
// Every path owns one (pixel, wavelength) slot, so there are no floating-point atomics and the image is
// deterministic.
#include <cuda_runtime.h>

#define FT_WITH_CONE 1
#define FT_SCAN_UNROLL 2  // these kernels are larger and instruction-cache sensitive: four records per round measured slower here
#define FT_NS ft_spectral
#include "kernels.h"
#include "philox.h"
#include "spectral.h"
#include "trace_device.cuh"

namespace ft_spectral {

constexpr int SB = 128;  // threads per CTA
constexpr unsigned FULL = 0xffffffffu;
#ifndef FT_SPEC_REFILL_MIN
#define FT_SPEC_REFILL_MIN 32  // idle lanes of a path-wave warp that trigger a refill (32: whole chunks in lockstep).  Measured on
                               // B200 (config 3): 1, 8, 16, 24 are all 5-8 % slower than 32 -- the chunk cursor costs registers in a
                               // kernel that already spills, and refilled lanes drift apart
#endif
#ifndef FT_PROBE_INLINE
#define FT_PROBE_INLINE 0  // 1: the once-per-warp probe of the camera position is a second inlined copy of the evaluator
#endif
#ifndef FT_SPEC_MINB
#define FT_SPEC_MINB 7  // <= 72 registers.  Measured (config 3, 4K; round 2, sub-grid + batched material stage): 6 CTAs/SM 60.2 ms, 7: 58.1, 8: 59.2
#endif

enum : int {
    SP_PARK = -1,   // path waves: a glass hit waiting for the warp's batched material stage (see the fetch block)
    SP_FETCH = 0,
    SP_MARCH = 1,   // primary ray or continuation path
    SP_SHADOW = 2,  // the reference's shadow ray (SdfLight.fs:10-20 / 26-41)
    SP_N0 = 3,      // normal probes +x, +y, +z, centre (SdfForm.fs:106-112)
    SP_N1 = 4,
    SP_N2 = 5,
    SP_N3 = 6,
};

// ---- spectrum: the frequency bands of Light.Colors (Light.fs:21-26), THz
__device__ __forceinline__ int band_of(float nu) {
    return nu < 484.0f ? 0 : nu < 508.0f ? 1 : nu < 526.0f ? 2 : nu < 606.0f ? 3 : nu < 668.0f ? 4 : 5;
}
__device__ __forceinline__ void band_weight(int b, float &w0, float &w1, float &w2) {
    // red, orange, yellow, green, blue, violet
    w0 = b <= 2 ? 1.0f : b == 5 ? 0.5f : 0.0f;
    w1 = b == 1 ? 0.5f : (b == 2 || b == 3) ? 1.0f : 0.0f;
    w2 = b >= 4 ? 1.0f : 0.0f;
}
__device__ __forceinline__ float spec(V3 rgb, int b) {
    float w0, w1, w2;
    band_weight(b, w0, w1, w2);
    return ((rgb.x * w0 + rgb.y * w1) + rgb.z * w2) / ((w0 + w1) + w2);
}
__device__ __forceinline__ float wavelength_nu(int k, int n) { return 400.0f + 389.0f * (((float)k + 0.5f) / (float)n); }

__device__ __forceinline__ float pick6(const float (&a)[6], int b) {
    float v = a[0];
#pragma unroll
    for (int i = 1; i < 6; i++)
        if (b == i) v = a[i];
    return v;
}

// pixel RGB = sum_k w(nu_k) L_k / sum_k w(nu_k), accumulated in k order; Lb[b] is L for the band of nu_k
__device__ __forceinline__ void write_pixel_from_bands(const StripTarget &T, long long pixel, int nl, const float (&Lb)[6]) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int k = 0; k < nl; k++) {
        const int b = band_of(wavelength_nu(k, nl));
        float w0, w1, w2;
        band_weight(b, w0, w1, w2);
        const float L = pick6(Lb, b);
        a0 = a0 + w0 * L; a1 = a1 + w1 * L; a2 = a2 + w2 * L;
        d0 = d0 + w0; d1 = d1 + w1; d2 = d2 + w2;
    }
    T.store((unsigned long long)pixel, d0 > 0.0f ? a0 / d0 : 0.0f, d1 > 0.0f ? a1 / d1 : 0.0f, d2 > 0.0f ? a2 / d2 : 0.0f);
}

// ---- the glass interaction: Light.fresnel n1 n2 N direction (Light.fs:30-59) as written, N = the surface normal
// oriented along the ray, plus the Philox choice between Reflect and Transmit.  Returns the continuation ray.
struct Continuation {
    V3 o, d;
    float len;
    bool into_glass;  // the medium the continuation is expected to run in (the march itself decides from the sign of Distance)
};
__device__ __forceinline__ Continuation glass_interact(V3 ab, float nu, uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel,
                                                       uint32_t k, uint32_t bounce, V3 ho, V3 hd, float hlen, float eps,
                                                       V3 normal, int inside) {
    const float lam = 299.792458f / nu;  // micrometres
    const float nn = ab.x + ab.y / (lam * lam);  // Cauchy
    const float n1 = inside ? nn : 1.0f, n2 = inside ? 1.0f : nn;
    const V3 N = dot(normal, hd) > 0.0f ? normal : -normal;
    const float n1n2 = n1 / n2;                               // Light.fs:37
    const float cosi = dot(N, hd);                            // Light.fs:38
    const float n1n2cosi = n1n2 * cosi;                       // Light.fs:39
    const float cost = sqrtf(1.0f + n1n2cosi * cosi - n1n2);  // Light.fs:40 (as written)
    float reflectance = 1.0f;                                 // NaN cost: total internal reflection
    if (!isnan(cost)) {
        float a = n2 * cosi, b = n1 * cost;
        float x = (a - b) / (a + b);
        const float rs = x * x;  // Light.fs:43-47
        a = n1 * cosi;
        b = n2 * cost;
        x = (a - b) / (a + b);
        const float rp = x * x;  // Light.fs:49-53
        reflectance = 0.5f * (rs + rp);
    }
    uint32_t c[4] = {pixel, k, bounce, 0u};
    ft_philox4x32_10(c, seed_lo, seed_hi);
    const float u = __uint2float_rn(c[0]) * 2.3283064365386963e-10f;  // == (float)(c0 * 2^-32 in double)
    Continuation r;
    const bool reflect = u < reflectance;
    if (reflect) r.d = normalize(hd - N * (2.0f * cosi));     // Light.fs:57
    else r.d = normalize(hd * n1n2 + N * (n1n2cosi - cost));  // Light.fs:58
    r.into_glass = reflect ? inside != 0 : inside == 0;
    r.o = ho + r.d * (2.0f * eps);
    r.len = hlen - 2.0f * eps;
    return r;
}

// queue append of the lanes with `push` set: ballot + popc ranks, one atomic per warp and class; returns the lane's index
__device__ __forceinline__ unsigned queue_slot(const SpectralParams &P, bool push, bool into_glass, int lane) {
    const bool front = into_glass || !P.split;
    const unsigned pa = __ballot_sync(FULL, push && front), pb = __ballot_sync(FULL, push && !front);
    unsigned base_a = 0, base_b = 0;
    if (pa) {
        const int leader = __ffs(pa) - 1;
        if (lane == leader) base_a = atomicAdd(P.out_count, (unsigned)__popc(pa));
        base_a = __shfl_sync(FULL, base_a, leader);
    }
    if (pb) {
        const int leader = __ffs(pb) - 1;
        if (lane == leader) base_b = atomicAdd(P.out_count + 1, (unsigned)__popc(pb));
        base_b = __shfl_sync(FULL, base_b, leader);
    }
    const unsigned below = (1u << lane) - 1u;
    return front ? base_a + (unsigned)__popc(pa & below) : P.q_cap - 1u - (base_b + (unsigned)__popc(pb & below));
}

__device__ __forceinline__ void store_path(float4 *q, unsigned idx, const Continuation &c, uint32_t pixel, uint32_t slot, uint32_t k) {
    q[3ull * idx + 0] = float4{c.o.x, c.o.y, c.o.z, c.d.x};
    q[3ull * idx + 1] = float4{c.d.y, c.d.z, c.len, __uint_as_float(pixel)};
    q[3ull * idx + 2] = float4{__uint_as_float(slot), __uint_as_float(k), 0.f, 0.f};
}

// ---- wave kernels ----------------------------------------------------------------------------------------
// One distance-evaluation call site; every lane runs a small state machine (march -> 4 normal probes -> light
// loop with shadow marches), whole 32-item chunks in lockstep like the hot path's ray machine.
// algorithmic float32 operations of the extension's own steps (the evaluator counts its own, trace_device.cuh)
constexpr int FLOPS_SPEC = 8;            // spec(rgb, band): 3 mul, 4 add, 1 div
constexpr int FLOPS_GLASS = 60;          // Cauchy index, Light.fresnel, reflect / transmit direction, new origin
constexpr int FLOPS_BAND_SUM = 9;        // per wavelength of the pixel sum: 3 mul, 6 add

// COUNT: instrumented variant (algorithmic flops and march steps into the stats block; bench.py's roofline numerator)
template <bool CHAIN, bool PRIMARY, bool COUNT>
__global__ void __launch_bounds__(SB, FT_SPEC_MINB) spectral_trace_kernel(const __grid_constant__ DevScene S, const __grid_constant__ SpectralParams P) {
    const int lane = threadIdx.x & 31;
    const float eps = P.epsilon;
    // path waves: the front class occupies whole chunks [0, front_pad), the back class follows
    const unsigned n_front = PRIMARY ? 0u : __ldg(P.in_count), n_back = PRIMARY ? 0u : __ldg(P.in_count + 1);
    const long long front_pad = ((long long)n_front + 31) & ~31ll;
    const long long total = PRIMARY ? P.n_items : front_pad + (long long)n_back;
    constexpr int NB = PRIMARY ? 6 : 1;

    Acc acc{0ull, 0ull};
    uint32_t hit_best = 0u;  // eval_chain's arg-min note: kept at a hit, it names the material's leaf (FT_CHAIN_MAT)
    float d0 = 0.f;
    if (PRIMARY) {  // Distance at the common origin of all primary rays, once per warp
        const V3 c = V3{P.cam.position[0], P.cam.position[1], P.cam.position[2]};
        #if FT_PROBE_INLINE
        uint32_t best_dummy = 0u;
        d0 = CHAIN ? eval_chain<COUNT>(S, c, acc, best_dummy) : eval_form<COUNT>(S, S.root_word, c, acc);
#else
        d0 = CHAIN ? eval_chain_call<COUNT>(S, c, acc) : eval_form_call<COUNT>(S, S.root_word, c, acc);
#endif
    }

    int phase = SP_FETCH;
#if FT_SPEC_REFILL_MIN < 32
    long long chunk_next = 0;  // the warp's current chunk (uniform)
    int chunk_left = 0;
    bool exhausted = false;
#endif
    // (o, d, len) stay the un-moved hit ray from the hit until the material is known: the normal probes evaluate around q and
    // leave them alone, so no copy of the hit ray is kept
    V3 o{0, 0, 0}, d{0, 0, 0}, q{0, 0, 0}, n{0, 0, 0};
    float len = 0.f, lcb[NB];
    float sign = 1.0f;
    bool armed = true;
    int steps = 0, light = 0, inside = 0, band = 0, mat = 0;
    uint32_t pixel = 0, slot = 0, kk = 0;
    unsigned n_hits = 0, n_shadow = 0, n_capped = 0, n_escaped = 0, n_paths = 0;
#pragma unroll
    for (int b = 0; b < NB; b++) lcb[b] = 0.f;

    for (;;) {
#if FT_SPEC_REFILL_MIN >= 32
        // ------------------------------------------------------------------ a new chunk once the whole warp is idle
        if (__all_sync(FULL, phase <= SP_FETCH)) {
            // Path waves: the material stage of the chunk's glass hits runs here, for all of them at once.  Done in each lane's
            // own transition it ran a dozen times per chunk with a few lanes each (Fresnel + Philox + the queue append are ~250
            // instructions: a quarter of the wave's warp instructions at 3-6 active lanes); the lanes would have idled anyway.
            if (!PRIMARY && __any_sync(FULL, phase == SP_PARK)) {
                const bool parked = phase == SP_PARK;
                Continuation cont{V3{0, 0, 0}, V3{0, 0, 0}, 0.f, false};
                if (parked) {
                    cont = glass_interact(xyz(ld4(S.materials + mat)), wavelength_nu((int)kk, P.nl), P.seed_lo, P.seed_hi, pixel, kk,
                                          (uint32_t)P.wave, o, d, len, eps, n, inside);
                    add_flops<COUNT>(acc, FLOPS_GLASS);
                }
                const unsigned idx = queue_slot(P, parked, cont.into_glass, lane);
                if (parked) {
                    store_path(P.out_q, idx, cont, pixel, slot, kk);
                    phase = SP_FETCH;
                }
            }
            // Wave 0: the tile's glass hits enter the hit queue together, in tile order -- the material stage and the path waves
            // take 32 consecutive entries per warp, and appended hit by hit (as each lane finished its normal probes) the queue
            // interleaved the pixels of all the tiles in flight: a path-wave chunk held two unrelated pixels and waited for the
            // longer of their segments
            if (PRIMARY && __any_sync(FULL, phase == SP_PARK)) {
                const bool parked = phase == SP_PARK;
                const unsigned pm = __ballot_sync(FULL, parked);
                const int leader = __ffs(pm) - 1;
                unsigned qbase = 0;
                if (lane == leader) qbase = atomicAdd(P.hit_count, (unsigned)__popc(pm));
                qbase = __shfl_sync(FULL, qbase, leader);
                if (parked) {
                    const unsigned idx = qbase + (unsigned)__popc(pm & ((1u << lane) - 1u));
                    P.hit_q[3ull * idx + 0] = float4{o.x, o.y, o.z, len};
                    P.hit_q[3ull * idx + 1] = float4{d.x, d.y, d.z, __uint_as_float(pixel)};
                    P.hit_q[3ull * idx + 2] = float4{n.x, n.y, n.z, __int_as_float(mat)};
                    phase = SP_FETCH;
                }
            }
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.work_counter, 32ull);
            base = __shfl_sync(FULL, base, 0);
            if ((long long)base >= total) break;
            const long long i = (long long)base + lane;
            if (i < total && (PRIMARY || i >= front_pad || i < (long long)n_front)) {
                if (PRIMARY) {  // 4 (x) by 8 (y) pixel tiles, like Image.render's machine
                    const int tiles_y = (P.height + 7) / 8;
                    const bool by_tile = P.T.strip_width == 0;  // tiles dealt round-robin over the ranks instead of column strips
                    const long long tile = by_tile ? (i >> 5) * P.T.world + P.T.rank : (i >> 5);
                    const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
                    const int xl = tx * 4 + (lane >> 3), y = ty * 8 + (lane & 7);  // xl: column of the rank's compact strip space
                    const int x = by_tile ? xl : P.T.image_column(xl);
                    if ((by_tile || xl < P.T.local_cols) && x < P.width && y < P.height) {
                        pixel = (uint32_t)((long long)x * P.height + y);
                        o = V3{P.cam.position[0], P.cam.position[1], P.cam.position[2]};
                        d = pixel_direction(P.cam, x, y, P.max_size);
                        len = P.length;
                        steps = 0;
                        if (len > 0.0f && d0 >= eps) {  // first iteration of the march with the shared Distance
                            o = o + d * d0;
                            len = len - d0;
                            steps = 1;
                        }
                        sign = 1.0f;
                        armed = true;
                        inside = 0;
                        phase = SP_MARCH;
                        add_flops<COUNT>(acc, FLOPS_PRIMARY_RAY);
                    }
                } else {
                    const long long qi = i < front_pad ? i : (long long)P.q_cap - 1 - (i - front_pad);
                    const float4 a = __ldg(P.in_q + 3 * qi), b = __ldg(P.in_q + 3 * qi + 1), c = __ldg(P.in_q + 3 * qi + 2);
                    o = V3{a.x, a.y, a.z};
                    d = V3{a.w, b.x, b.y};
                    len = b.z;
                    pixel = __float_as_uint(b.w);
                    slot = __float_as_uint(c.x);
                    kk = __float_as_uint(c.y);
                    band = band_of(wavelength_nu((int)kk, P.nl));
                    sign = 1.0f;
                    inside = 0;
                    armed = false;  // continuation rays arm (and learn their medium) once |Distance| >= epsilon
                    steps = 0;
                    n_paths++;
                    phase = SP_MARCH;
                }
            }
            continue;
        }
#else
        // ------------------------------------------------------------------ refill
        // Wave 0 takes a new 4 x 8 pixel tile once the whole warp is idle (coherent primary rays: lockstep pays, see the hot path).
        // Path waves refill as soon as FT_SPEC_REFILL_MIN lanes are idle: the segments of one chunk differ several-fold in length
        // (grazing paths crawl along the glass surface), and consecutive queue entries are the other wavelengths of the same
        // pixels, so a refilled lane stays close to its neighbours.
        {
            unsigned idle = __ballot_sync(FULL, phase == SP_FETCH);
            constexpr int REFILL = PRIMARY ? 32 : FT_SPEC_REFILL_MIN;
            if (idle == FULL || (__popc(idle) >= REFILL && !(exhausted && chunk_left == 0))) {
                while (idle) {
                    if (chunk_left == 0) {
                        if (exhausted) break;
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(P.work_counter, 32ull);
                        base = __shfl_sync(FULL, base, 0);
                        if ((long long)base >= total) {
                            exhausted = true;
                            break;
                        }
                        chunk_next = (long long)base;
                        chunk_left = 32;
                    }
                    const int want = __popc(idle), take = want < chunk_left ? want : chunk_left;
                    const int r = __popc(idle & ((1u << lane) - 1u));
                    const long long i = chunk_next + r;
                    if (phase == SP_FETCH && r < take && i < total && (PRIMARY || i >= front_pad || i < (long long)n_front)) {
                        if (PRIMARY) {  // 4 (x) by 8 (y) pixel tiles, like Image.render's machine (take == 32: r == lane)
                            const int tiles_y = (P.height + 7) / 8;
                            const bool by_tile = P.T.strip_width == 0;  // tiles dealt round-robin over the ranks instead of column strips
                            const long long tile = by_tile ? (i >> 5) * P.T.world + P.T.rank : (i >> 5);
                            const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
                            const int xl = tx * 4 + (r >> 3), y = ty * 8 + (r & 7);  // xl: column of the rank's compact strip space
                            const int x = by_tile ? xl : P.T.image_column(xl);
                            if ((by_tile || xl < P.T.local_cols) && x < P.width && y < P.height) {
                                pixel = (uint32_t)((long long)x * P.height + y);
                                o = V3{P.cam.position[0], P.cam.position[1], P.cam.position[2]};
                                d = pixel_direction(P.cam, x, y, P.max_size);
                                len = P.length;
                                steps = 0;
                                if (len > 0.0f && d0 >= eps) {  // first iteration of the march with the shared Distance
                                    o = o + d * d0;
                                    len = len - d0;
                                    steps = 1;
                                }
                                sign = 1.0f;
                                armed = true;
                                inside = 0;
                                phase = SP_MARCH;
                                add_flops<COUNT>(acc, FLOPS_PRIMARY_RAY);
                            }
                        } else {
                            const long long qi = i < front_pad ? i : (long long)P.q_cap - 1 - (i - front_pad);
                            const float4 a = __ldg(P.in_q + 3 * qi), b = __ldg(P.in_q + 3 * qi + 1), c = __ldg(P.in_q + 3 * qi + 2);
                            o = V3{a.x, a.y, a.z};
                            d = V3{a.w, b.x, b.y};
                            len = b.z;
                            pixel = __float_as_uint(b.w);
                            slot = __float_as_uint(c.x);
                            kk = __float_as_uint(c.y);
                            band = band_of(wavelength_nu((int)kk, P.nl));
                            sign = 1.0f;
                            inside = 0;
                            armed = false;  // continuation rays arm (and learn their medium) once |Distance| >= epsilon
                            steps = 0;
                            n_paths++;
                            phase = SP_MARCH;
                        }
                    }
                    chunk_next += take;
                    chunk_left -= take;
                    idle = __ballot_sync(FULL, phase == SP_FETCH);  // padding items leave their lane idle: it draws again ...
                    if (PRIMARY && idle != FULL) break;             // ... but a warp of wave 0 never mixes two tiles
                }
                if (idle == FULL) {
                    if (exhausted) break;  // no work left anywhere
                    continue;
                }
            }
        }
#endif
        // ------------------------------------------------------------------ march pre-check
        int event = 0;  // 1 march miss, 2 march hit, 3 shadow ray unoccluded, 4 shadow ray occluded, 5 normal done
        V3 p = o;  // the evaluation point
        if (phase >= SP_N0) {
            const float h = eps * 0.125f;
            p = q;
            if (phase == SP_N0) p.x = q.x + h;
            else if (phase == SP_N1) p.y = q.y + h;
            else if (phase == SP_N2) p.z = q.z + h;
        }
        // chain scenes: one box test on the cell coordinate decides "no escape pre-check" and "straight to the sub-grid"
        const V3 cc = CHAIN ? chain_coord(S, p) : V3{0.f, 0.f, 0.f};
        const bool fast = CHAIN && in_fast_region(S, cc);
        if (phase == SP_MARCH || phase == SP_SHADOW) {
            if (len <= 0.0f || steps >= S.step_cap) {
                if (!(len <= 0.0f)) n_capped++;
                event = phase == SP_MARCH ? 1 : 3;
            } else if (S.esc_valid && !fast && (phase == SP_SHADOW || !armed || sign > 0.0f) && outside_box(S.esc_lo, S.esc_hi, o) &&
                       ray_escapes(S, o, d, eps)) {
                n_escaped++;  // Distance >= distance to the escape box > epsilon along the whole half-line: no hit ahead
                event = phase == SP_MARCH ? 1 : 3;
            }
        }
        // ------------------------------------------------------------------ the one distance evaluation
        if (phase > SP_FETCH && event == 0) {
            uint32_t best_lo = 0u;
            const float dist = CHAIN ? eval_chain_c<COUNT>(S, p, cc, fast, acc, best_lo) : eval_form<COUNT>(S, S.root_word, p, acc);
            if (COUNT && phase <= SP_SHADOW) {
                acc.steps++;
                add_flops<COUNT>(acc, FLOPS_MOVE + 2);
            }
            if (phase == SP_MARCH) {
                if (!armed) sign = dist < 0.0f ? -1.0f : 1.0f;  // the medium: inside glass the march runs on -Distance
                const float sd = sign * dist;
                steps++;
                if (armed) {
                    if (sd < eps) {
                        event = 2;
                        hit_best = best_lo;  // the hit evaluation's arg-min candidate
                    } else {
                        o = o + d * sd;
                        len = len - sd;
                    }
                } else if (sd >= eps) {
                    armed = true;
                    o = o + d * sd;
                    len = len - sd;
                } else {
                    o = o + d * eps;
                    len = len - eps;
                }
            } else if (phase == SP_SHADOW) {
                steps++;
                if (dist < eps) {
                    event = 4;
                } else {
                    o = o + d * dist;
                    len = len - dist;
                }
            } else if (phase == SP_N0) {
                n.x = dist;
                phase = SP_N1;
            } else if (phase == SP_N1) {
                n.y = dist;
                phase = SP_N2;
            } else if (phase == SP_N2) {
                n.z = dist;
                phase = SP_N3;
            } else {
                n = normalize(n - V3{dist, dist, dist});
                add_flops<COUNT>(acc, FLOPS_NORMAL);
                event = 5;
            }
        }
        // ------------------------------------------------------------------ transitions
        bool push = false;
        Continuation cont{V3{0, 0, 0}, V3{0, 0, 0}, 0.f, false};
        if (event != 0) {
            bool next_light = false;
            if (event == 1) {  // miss: spec(BackgroundColor, nu)
                const V3 bg = V3{S.bg[0], S.bg[1], S.bg[2]};
                if (PRIMARY) {
                    float Lb[6];
#pragma unroll
                    for (int b = 0; b < 6; b++) Lb[b] = spec(bg, b);
                    write_pixel_from_bands(P.T, pixel, P.nl, Lb);
                    add_flops<COUNT>(acc, 6 * FLOPS_SPEC + FLOPS_BAND_SUM * P.nl + 3);
                } else {
                    P.slots[slot] = spec(bg, band);
                    add_flops<COUNT>(acc, FLOPS_SPEC);
                }
                phase = SP_FETCH;
            } else if (event == 2) {  // hit: keep the un-moved hit ray, probe the normal
                inside = sign < 0.0f;
                q = o + d * (-eps);
                n_hits++;
                phase = SP_N0;
            } else if (event == 5) {  // normal done: Material arg-min at the un-moved hit origin (SdfObject.fs:75-77)
                const int leaf_obj = (FT_CHAIN_MAT && CHAIN && S.chain_mat_valid) ? chain_leaf_from_best(S, o, hit_best)
                                                                                  : material_leaf<COUNT>(S, o, acc);
                mat = __ldg(S.objects + leaf_obj).z;
                const float4 m = ld4(S.materials + mat);
                if (__float_as_uint(m.w) == (uint32_t)FT_MATERIAL_GLASS) {
#if FT_SPEC_REFILL_MIN >= 32
                    if (PRIMARY) {
                        phase = SP_PARK;  // the tile's glass hits enter the hit queue together (fetch block)
                    } else if (P.wave >= P.max_bounces) {
                        P.slots[slot] = 0.0f;  // interaction budget spent
                    } else {
                        phase = SP_PARK;  // the material stage runs for the whole chunk at once (fetch block)
                    }
                    if (phase != SP_PARK) phase = SP_FETCH;
#else
                    if (PRIMARY) {
                        push = true;  // the material stage takes it from the hit queue
                    } else if (P.wave >= P.max_bounces) {
                        P.slots[slot] = 0.0f;  // interaction budget spent
                    } else {
                        cont = glass_interact(xyz(m), wavelength_nu((int)kk, P.nl), P.seed_lo, P.seed_hi, pixel, kk,
                                              (uint32_t)P.wave, o, d, len, eps, n, inside);
                        add_flops<COUNT>(acc, FLOPS_GLASS);
                        push = true;
                    }
                    phase = SP_FETCH;
#endif
                } else {
                    const V3 bg = V3{S.bg[0], S.bg[1], S.bg[2]};  // SdfScene.fs:12
#pragma unroll
                    for (int b = 0; b < NB; b++) lcb[b] = spec(bg, PRIMARY ? b : band);
                    light = 0;
                    next_light = true;
                }
            } else {  // shadow ray finished
                if (event == 3) {
                    // the light's intensity and cosine are formed again from its record (same inputs, same operations as when
                    // the shadow ray was set up) instead of being carried in four registers through the shadow march
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    V3 cint = xyz(lrgb);
                    float lcos;
                    if (__float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL) {
                        lcos = dot(n, xyz(lv));
                    } else {
                        lcos = dot(n, normalize(xyz(lv) - q));
                        cint = xyz(lrgb) / length2(xyz(lv) - q);
                    }
#pragma unroll
                    for (int b = 0; b < NB; b++) lcb[b] = lcb[b] + spec(cint, PRIMARY ? b : band) * lcos;
                    add_flops<COUNT>(acc, NB * (FLOPS_SPEC + 2));
                }
                light++;
                next_light = true;
            }
            if (next_light) {
                bool finish = true;
                for (; light < S.n_lights; light++) {  // SdfScene.fs:13-19
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    const bool directional = __float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL;
                    const V3 ldir = directional ? xyz(lv) : normalize(xyz(lv) - q);
                    const float lcos = dot(n, ldir);
                    add_flops<COUNT>(acc, FLOPS_LIGHT);
                    if (lcos > 0.0f) {
                        o = q;
                        V3 cint = xyz(lrgb);
                        if (directional) {
                            d = xyz(lv);
                            len = 1000.0f;
                        } else {
                            const V3 diff = xyz(lv) - q;
                            const float d2 = length2(diff);
                            len = sqrtf(d2);
                            d = diff / d2;
                            cint = xyz(lrgb) / d2;
                            add_flops<COUNT>(acc, FLOPS_POINT_LIGHT);
                        }
                        if (inside) {  // inside glass no shadow ray is cast: the light contributes
#pragma unroll
                            for (int b = 0; b < NB; b++) lcb[b] = lcb[b] + spec(cint, PRIMARY ? b : band) * lcos;
                            continue;
                        }
                        steps = 0;
                        n_shadow++;
                        phase = SP_SHADOW;
                        finish = false;
                        break;
                    }
                }
                if (finish) {  // SdfScene.fs:28 per wavelength
                    const V3 albedo = xyz(ld4(S.materials + mat));
                    const float pi_inv = 1.0f / 3.14159274f;
                    if (PRIMARY) {
                        float Lb[6];
#pragma unroll
                        for (int b = 0; b < 6; b++) Lb[b] = spec(albedo, b) * (lcb[b < NB ? b : 0] * pi_inv);
                        write_pixel_from_bands(P.T, pixel, P.nl, Lb);
                        add_flops<COUNT>(acc, 6 * (FLOPS_SPEC + 2) + FLOPS_BAND_SUM * P.nl + 3);
                    } else {
                        P.slots[slot] = spec(albedo, band) * (lcb[0] * pi_inv);
                        add_flops<COUNT>(acc, FLOPS_SPEC + 2);
                    }
                    phase = SP_FETCH;
                }
            }
        }
        // ------------------------------------------------------------------ queue append: ballot + popc ranks, one atomic per warp
        if (PRIMARY) {
            const unsigned pm = __ballot_sync(FULL, push);
            if (pm) {
                const int leader = __ffs(pm) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(P.hit_count, (unsigned)__popc(pm));
                base = __shfl_sync(FULL, base, leader);
                if (push) {
                    const unsigned idx = base + (unsigned)__popc(pm & ((1u << lane) - 1u));
                    P.hit_q[3ull * idx + 0] = float4{o.x, o.y, o.z, len};
                    P.hit_q[3ull * idx + 1] = float4{d.x, d.y, d.z, __uint_as_float(pixel)};
                    P.hit_q[3ull * idx + 2] = float4{n.x, n.y, n.z, __int_as_float(mat)};
                }
            }
        } else if (__any_sync(FULL, push)) {
            const unsigned idx = queue_slot(P, push, cont.into_glass, lane);
            if (push) store_path(P.out_q, idx, cont, pixel, slot, kk);
        }
    }

    const unsigned hits = __reduce_add_sync(FULL, n_hits), shadows = __reduce_add_sync(FULL, n_shadow);
    const unsigned capped = __reduce_add_sync(FULL, n_capped), escaped = __reduce_add_sync(FULL, n_escaped);
    const unsigned paths = __reduce_add_sync(FULL, n_paths);
    if (lane == 0) {
        const unsigned w = (blockIdx.x * (SB / 32) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1);
        unsigned long long *s = P.stats + w * FT_STAT_STRIDE;
        if (PRIMARY && hits) atomicAdd(s + FT_STAT_HITS, (unsigned long long)hits);
        if (shadows) atomicAdd(s + FT_STAT_SHADOW, (unsigned long long)shadows);
        if (capped) atomicAdd(s + FT_STAT_CAPPED, (unsigned long long)capped);
        if (escaped) atomicAdd(s + FT_STAT_ESCAPED, (unsigned long long)escaped);
        if (paths) atomicAdd(s + FT_STAT_PATHS, (unsigned long long)paths);
    }
    if (COUNT) {
        unsigned long long fl = acc.flops, st = acc.steps;
        for (int k = 16; k > 0; k >>= 1) {
            fl += __shfl_xor_sync(FULL, fl, k);
            st += __shfl_xor_sync(FULL, st, k);
        }
        if (lane == 0) {
            const unsigned w = (blockIdx.x * (SB / 32) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1);
            atomicAdd(P.stats + w * FT_STAT_STRIDE + FT_STAT_FLOPS, fl);
            atomicAdd(P.stats + w * FT_STAT_STRIDE + FT_STAT_STEPS, st);
        }
    }
}

// ---- the material stage of wave 0: one thread per (glass hit, wavelength); the continuations enter wave 1's queue by class
__global__ void __launch_bounds__(SB) spectral_spawn_kernel(const __grid_constant__ DevScene S, const __grid_constant__ SpectralParams P,
                                                            unsigned n_hits) {
    const unsigned long long t = (unsigned long long)blockIdx.x * SB + threadIdx.x;
    const bool live = t < (unsigned long long)n_hits * (unsigned)P.nl;
    if (P.max_bounces < 1) {  // no interaction allowed: the path carries nothing
        if (live) P.slots[t] = 0.0f;
        return;
    }
    Continuation cont{V3{0, 0, 0}, V3{0, 0, 0}, 0.f, false};
    uint32_t pixel = 0u, k = 0u;
    if (live) {
        const unsigned h = (unsigned)(t / (unsigned)P.nl);
        k = (unsigned)(t - (unsigned long long)h * (unsigned)P.nl);
        const float4 a = __ldg(P.hit_q + 3ull * h), b = __ldg(P.hit_q + 3ull * h + 1), c = __ldg(P.hit_q + 3ull * h + 2);
        const float4 m = ld4(S.materials + __float_as_int(c.w));
        pixel = __float_as_uint(b.w);
        cont = glass_interact(xyz(m), wavelength_nu((int)k, P.nl), P.seed_lo, P.seed_hi, pixel, k, 0u, V3{a.x, a.y, a.z},
                              V3{b.x, b.y, b.z}, a.w, P.epsilon, V3{c.x, c.y, c.z}, 0);
    }
    const unsigned idx = queue_slot(P, live, cont.into_glass, (int)(threadIdx.x & 31));
    if (live) store_path(P.out_q, idx, cont, pixel, (unsigned)t, k);
}

// ---- per glass pixel: RGB = sum_k w(nu_k) L_k / sum_k w(nu_k) in k order
__global__ void __launch_bounds__(SB) spectral_resolve_kernel(const __grid_constant__ SpectralParams P, unsigned n_hits) {
    const unsigned h = blockIdx.x * SB + threadIdx.x;
    if (h >= n_hits) return;
    const uint32_t pixel = __float_as_uint(__ldg(P.hit_q + 3ull * h + 1).w);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int k = 0; k < P.nl; k++) {
        float w0, w1, w2;
        band_weight(band_of(wavelength_nu(k, P.nl)), w0, w1, w2);
        const float L = P.slots[(unsigned long long)h * (unsigned)P.nl + k];
        a0 = a0 + w0 * L; a1 = a1 + w1 * L; a2 = a2 + w2 * L;
        d0 = d0 + w0; d1 = d1 + w1; d2 = d2 + w2;
    }
    P.T.store(pixel, d0 > 0.0f ? a0 / d0 : 0.0f, d1 > 0.0f ? a1 / d1 : 0.0f, d2 > 0.0f ? a2 / d2 : 0.0f);
}

template <class K>
static cudaError_t wave_grid(K kernel, long long items, int sm_count, unsigned *grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SB, 0);
    if (e != cudaSuccess) return e;
    long long g = (long long)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const long long need = (items + SB - 1) / SB;
    if (g > need) g = need;
    *grid = (unsigned)(g < 1 ? 1 : g);
    return cudaSuccess;
}

}  // namespace ft_spectral

#define FT_SPEC_LAUNCH(CHAIN, PRIM, COUNT, ITEMS)                                                                       \
    do {                                                                                                              \
        if ((e = wave_grid(spectral_trace_kernel<CHAIN, PRIM, COUNT>, ITEMS, sm_count, &grid)) != cudaSuccess) return e; \
        spectral_trace_kernel<CHAIN, PRIM, COUNT><<<grid, SB, 0, stream>>>(S, P);                                      \
    } while (0)
template <bool PRIM>
static cudaError_t spectral_launch_trace(const DevScene &S, const SpectralParams &P, long long items, bool count, int sm_count,
                                         cudaStream_t stream) {
    using namespace ft_spectral;
    unsigned grid = 1;
    cudaError_t e;
    if (S.chain_valid) {
        if (count) FT_SPEC_LAUNCH(true, PRIM, true, items);
        else FT_SPEC_LAUNCH(true, PRIM, false, items);
    } else {
        if (count) FT_SPEC_LAUNCH(false, PRIM, true, items);
        else FT_SPEC_LAUNCH(false, PRIM, false, items);
    }
    return cudaGetLastError();
}
#undef FT_SPEC_LAUNCH

cudaError_t spectral_launch_primary(const DevScene &S, const SpectralParams &P, bool count, int sm_count, cudaStream_t stream) {
    return spectral_launch_trace<true>(S, P, P.n_items, count, sm_count, stream);
}

cudaError_t spectral_launch_spawn(const DevScene &S, const SpectralParams &P, unsigned n_hits, cudaStream_t stream) {
    using namespace ft_spectral;
    const unsigned long long threads = (unsigned long long)n_hits * (unsigned)P.nl;
    if (threads == 0) return cudaSuccess;
    spectral_spawn_kernel<<<(unsigned)((threads + SB - 1) / SB), SB, 0, stream>>>(S, P, n_hits);
    return cudaGetLastError();
}

// max_items: an upper bound of the live paths of this wave (the true count is read from P.in_count on the device)
cudaError_t spectral_launch_wave(const DevScene &S, const SpectralParams &P, long long max_items, bool count, int sm_count,
                                 cudaStream_t stream) {
    if (max_items <= 0) return cudaSuccess;
    return spectral_launch_trace<false>(S, P, max_items, count, sm_count, stream);
}

cudaError_t spectral_launch_resolve(const SpectralParams &P, unsigned n_hits, cudaStream_t stream) {
    using namespace ft_spectral;
    if (n_hits == 0) return cudaSuccess;
    spectral_resolve_kernel<<<(n_hits + SB - 1) / SB, SB, 0, stream>>>(P, n_hits);
    return cudaGetLastError();
}
