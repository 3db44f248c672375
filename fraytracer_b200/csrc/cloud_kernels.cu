// cloud_kernels.cu -- EXTENSION (BASELINE config 4): volumetric cloud material, stochastic, N samples per pixel.
//
// NOT reference behaviour.  The reference has no participating media, no stochastic sampling and no
// anti-aliasing; the spec is this repository's own.  Reused unchanged from
// the reference: the march loop (SdfForm.fs:93-104), normalFromRay (SdfForm.fs:106-115), the material arg-min
// (SdfObject.fs:27-46), shadow rays and the Lambert sum for solid hits (SdfLight.fs:6-42, SdfScene.fs:12-28).
// The density field is the reference's dead-code lattice noise Noise.gradient (Noise.fs:78-110), float64, restated
// as written (including `z+fz`, Noise.fs:80); its tables arrive as data through the ABI.
// Parity target: oracle/ft_oracle.c fto_render_cloud, bit for bit (compiled with -fmad=false, IEEE div/sqrt;
// no transcendental function on the path: collisions are Bernoulli draws per fixed step, not exponential
// free paths, so that no libm call separates the two implementations).
//
// One persistent kernel.  A warp owns a 4 x 8 pixel tile; every lane runs ALL samples of its own pixel one after
// the other (per-lane refill costs no coherence here: consecutive samples of a pixel are near-identical rays) and
// sums them in sample order in registers -- deterministic, no atomics, no sample buffer.  As in the hot path's ray
// machine every lane is a state machine around ONE distance-evaluation call site, so lanes that march a camera
// ray, probe a normal, track through the medium or follow a shadow ray all execute the same evaluator
// instructions: the divergence this config stresses is confined to the (cheap) transition code.
#include <cuda_runtime.h>

#define FT_WITH_CONE 1
#define FT_SCAN_UNROLL 2  // these kernels are larger and instruction-cache sensitive: four records per round measured slower here
#define FT_NS ft_cloud
#include "cloud.h"
#include "kernels.h"
#include "philox.h"
#include "trace_device.cuh"

namespace ft_cloud {

constexpr int CB = 128;
#ifndef FT_CLOUD_REFILL_MIN
#define FT_CLOUD_REFILL_MIN 32  // lanes that must want their next sample before any of them starts it (1: at once; 32: the warp's
                                // lanes start their samples together).  Measured on B200 (config 4): 1: 264.2 ms, 4: 240.3, 8: 229.4,
                                // 16: 220.8, 24: 216.1, 32: 181.1 -- samples of neighbouring pixels started together march in step
#endif
#ifndef FT_CLOUD_MINB
#define FT_CLOUD_MINB 4  // <= 128 registers.  Measured (config 4, 16 spp): 2 or 3 CTAs/SM 94 ms, 4: 81.4, 6: 84.9, 8: 87.5
#endif
constexpr unsigned FULL = 0xffffffffu;

enum : int {
    C_FETCH = 0,
    C_MARCH = 1,   // camera ray segment (SdfForm.tryTrace)
    C_SHADOW = 2,  // the reference's shadow ray
    C_N0 = 3,
    C_N1 = 4,
    C_N2 = 5,
    C_N3 = 6,
    C_TRACK = 7,   // fixed-step tracking along the camera ray inside the medium
    C_LTRACK = 8,  // fixed-step tracking from the scatter point towards a light
};

struct NoiseTab {
    int perm[256];
    double grad[768];
};

__device__ __forceinline__ int nperm(const NoiseTab &n, int x) { return n.perm[x & 255]; }
__device__ __forceinline__ int nindex(const NoiseTab &n, int ix, int iy, int iz) { return nperm(n, ix + nperm(n, iy + nperm(n, iz))); }
__device__ __forceinline__ double glattice(const NoiseTab &n, int ix, int iy, int iz, double fx, double fy, double fz) {
    const double *g = n.grad + 3 * nindex(n, ix, iy, iz);
    return ((g[0] * fx + g[1] * fy) + g[2]) + fz;  // Noise.fs:80 as written
}
__device__ __forceinline__ double nlerp(double t, double x0, double x1) { return x0 + t * (x1 - x0); }
__device__ __forceinline__ double smooth_step(double x) { return (x * x) * (3.0 - 2.0 * x); }
__device__ __forceinline__ double lat_and_lerp(const NoiseTab &n, double w, int ix0, int iy0, int iz0, double fx0, double fy0,
                                               double fz0, int ix1, int iy1, int iz1, double fx1, double fy1, double fz1) {
    const double a = glattice(n, ix0, iy0, iz0, fx0, fy0, fz0);
    const double b = glattice(n, ix1, iy1, iz1, fx1, fy1, fz1);
    return nlerp(w, a, b);
}
__device__ double noise_gradient(const NoiseTab &n, double x, double y, double z) {  // Noise.fs:97-110
    const int ix = (int)floor(x), iy = (int)floor(y), iz = (int)floor(z);
    const double fx0 = x - (double)ix, fy0 = y - (double)iy, fz0 = z - (double)iz;
    const double fx1 = fx0 - 1.0, fy1 = fy0 - 1.0, fz1 = fz0 - 1.0;
    const double wx = smooth_step(fx0), wy = smooth_step(fy0), wz = smooth_step(fz0);
    double vy0 = lat_and_lerp(n, wx, ix, iy, iz, fx0, fy0, fz0, ix + 1, iy, iz, fx1, fy0, fz0);
    double vy1 = lat_and_lerp(n, wx, ix, iy + 1, iz, fx0, fy1, fz0, ix + 1, iy + 1, iz, fx1, fy1, fz0);
    const double vz0 = nlerp(wy, vy0, vy1);
    vy0 = lat_and_lerp(n, wx, ix, iy, iz + 1, fx0, fy0, fz1, ix + 1, iy, iz + 1, fx1, fy0, fz1);
    vy1 = lat_and_lerp(n, wx, ix, iy + 1, iz + 1, fx0, fy1, fz1, ix + 1, iy + 1, iz + 1, fx1, fy1, fz1);
    const double vz1 = nlerp(wy, vy0, vy1);
    return nlerp(wz, vz0, vz1);
}
__device__ __forceinline__ float clamp01f(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
__device__ __forceinline__ float cloud_density(const NoiseTab &n, const CloudParams &P, V3 p, float D) {
    const double f = (double)P.frequency;
    const float nz = (float)noise_gradient(n, (double)p.x * f, (double)p.y * f, (double)p.z * f);
    const float shape = clamp01f(P.bias + P.gain * nz);
    const float edge = clamp01f(-D / P.fade);
    return (P.density * shape) * edge;
}
// Random numbers: block b of stream `stream` of (pixel, sample) = Philox4x32-10({pixel, sample, b, stream}; seed).
// Stream 0 / block 0: the sub-pixel jitter (words 0, 1).  A tracking stream: block 0 word 0 = the collision draw of step 0,
// word 1 = the offset xi0 of the first step; the collision draw of step k >= 1 is word (k + 1) & 3 of block (k + 1) >> 2 -- one
// Philox evaluation serves four tracking steps (it was a fifth of a tracking step's instructions when every step drew its own).
__device__ __forceinline__ float cloud_u(uint32_t w) { return __uint2float_rn(w) * 2.3283064365386963e-10f; }  // == (float)(w * 2^-32 in double)
__device__ __noinline__ uint4 cloud_block(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t block,
                                          uint32_t stream) {  // one out-of-line copy: four call sites in a kernel that is
    uint32_t c[4] = {pixel, sample, block, stream};            // bound by instruction fetch
    ft_philox4x32_10(c, seed_lo, seed_hi);
    return make_uint4(c[0], c[1], c[2], c[3]);
}
// first draw of a stream: u0 = word 0, u1 = word 1; words 2 and 3 are parked in the thread's shared-memory slots for steps 1, 2
__device__ __forceinline__ void cloud_rnd(const CloudParams &P, uint32_t (*spare)[CB], uint32_t pixel, uint32_t sample, uint32_t stream,
                                          float &u0, float &u1) {
    const uint4 c = cloud_block(P.seed_lo, P.seed_hi, pixel, sample, 0u, stream);
    u0 = cloud_u(c.x);
    u1 = cloud_u(c.y);
    spare[1][threadIdx.x] = c.z;
    spare[2][threadIdx.x] = c.w;
}
// collision draw of tracking step k >= 1
__device__ __forceinline__ float cloud_track_u(const CloudParams &P, uint32_t (*spare)[CB], uint32_t pixel, uint32_t sample, uint32_t k,
                                               uint32_t stream) {
    const uint32_t j = k + 1u;
    if ((j & 3u) == 0u) {
        const uint4 c = cloud_block(P.seed_lo, P.seed_hi, pixel, sample, j >> 2, stream);
        spare[0][threadIdx.x] = c.y;
        spare[1][threadIdx.x] = c.z;
        spare[2][threadIdx.x] = c.w;
        return cloud_u(c.x);
    }
    return cloud_u(spare[(j & 3u) - 1u][threadIdx.x]);
}

// pixel of lane `lane` in local tile `tile`; false when it lies outside the frame (padding)
__device__ __forceinline__ bool cloud_tile_pixel(const CloudParams &P, int tiles_y, unsigned long long tile, int lane, int &x, int &y) {
    if (P.T.strip_width == 0) {  // tiles dealt round-robin over the ranks: the finest balance (cloud pixels cost ~50x)
        const unsigned long long gt = tile * (unsigned)P.T.world + (unsigned)P.T.rank;
        const int tx = (int)(gt / tiles_y), ty = (int)(gt - (unsigned long long)tx * tiles_y);
        x = tx * 4 + (lane >> 3);
        y = ty * 8 + (lane & 7);
        return x < P.width && y < P.height;
    }
    const int tx = (int)(tile / tiles_y), ty = (int)(tile - (unsigned long long)tx * tiles_y);
    const int xl = tx * 4 + (lane >> 3);  // column of the rank's compact strip space
    x = P.T.image_column(xl);
    y = ty * 8 + (lane & 7);
    return xl < P.T.local_cols && x < P.width && y < P.height;
}

// pixel = (((P0 + P1) + P2) + P3) / spp, the group sums in group order (FT_CLOUD_SAMPLE_GROUPS), stored to every destination
__global__ void __launch_bounds__(CB) cloud_resolve_kernel(const __grid_constant__ CloudParams P) {
    const long long i = (long long)blockIdx.x * CB + threadIdx.x;
    if (i >= P.n_tiles * 32) return;
    int x, y;
    if (!cloud_tile_pixel(P, (P.height + 7) / 8, (unsigned long long)(i >> 5), (int)(i & 31), x, y)) return;
    const float *pp = P.partials + i * FT_CLOUD_SAMPLE_GROUPS * 3;
    V3 t = V3{pp[0], pp[1], pp[2]};
#pragma unroll
    for (int g = 1; g < FT_CLOUD_SAMPLE_GROUPS; g++) t = t + V3{pp[3 * g], pp[3 * g + 1], pp[3 * g + 2]};
    const float inv = (float)P.spp;
    P.T.store((unsigned long long)x * P.height + y, t.x / inv, t.y / inv, t.z / inv);
}

// algorithmic operation counts of the extension's own steps (the evaluator counts its own, trace_device.cuh)
constexpr int FLOPS_TRACK_STEP = 11;  // t, p = o + d t, the two comparisons
constexpr int FLOPS_DENSITY = 95;     // cloud_density: Noise.gradient in float64 (82) + the float32 shaping (13)
constexpr int FLOPS_SAMPLE_RAY = 29;  // jittered camera ray (Camera.uniformPixelToRay + 4)

// COUNT: instrumented variant (algorithmic flops into the stats block; bench.py's roofline numerator)
template <bool CHAIN, bool COUNT>
__global__ void __launch_bounds__(CB, FT_CLOUD_MINB) cloud_kernel(const __grid_constant__ DevScene S, const __grid_constant__ CloudParams P) {
    __shared__ NoiseTab tab;
    __shared__ uint32_t spare[3][CB];  // the unused words of each thread's current Philox block (cloud_track_u)
    for (int i = threadIdx.x; i < 256; i += CB) tab.perm[i] = P.perm[i];
    for (int i = threadIdx.x; i < 768; i += CB) tab.grad[i] = P.gradients[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const float eps = P.epsilon;
    const uint32_t nstreams = 1u + (uint32_t)S.n_lights;
    const int tiles_y = (P.height + 7) / 8;
    const V3 cam = V3{P.cam.position[0], P.cam.position[1], P.cam.position[2]};
    const V3 bg = V3{S.bg[0], S.bg[1], S.bg[2]};
    Acc acc{0ull, 0ull};
    uint32_t hit_best = 0u;  // eval_chain's arg-min note: kept at a hit, it names the material's leaf (FT_CHAIN_MAT)
    // Distance at the common origin of all camera rays, once per warp
    const float d0 = CHAIN ? eval_chain_call<COUNT>(S, cam, acc) : eval_form_call<COUNT>(S, S.root_word, cam, acc);

    int phase = C_FETCH;
    bool have_tile = false, active = false, cloud_mode = false;
    int x = 0, y = 0, s = 0, spp_eff = 0, item_group = 0;  // spp_eff: one past the last sample of this work item
    long long item_tile = 0;
    uint32_t pixel = 0, seg = 0, stream = 0, k = 0;
    // (o, d, len) stay the un-moved hit ray while the normal is probed and while the camera ray is tracked through the medium
    // (neither touches them; d is the sample's camera direction until a shadow ray or a light track ends the sample), and a
    // light's intensity is formed again from its record when its visibility is decided: no copies are carried
    V3 sum{0, 0, 0}, o{0, 0, 0}, d{0, 0, 0}, q{0, 0, 0}, n{0, 0, 0}, ps{0, 0, 0}, lc{0, 0, 0};
    float len = 0.f, xi0 = 0.f, ucur = 0.f, maxdist = 0.f;
    int steps = 0, light = 0, mat = 0;
    unsigned n_hits = 0, n_secondary = 0, n_capped = 0, n_escaped = 0, n_scattered = 0;
    unsigned long long n_track = 0;

    for (;;) {
        // ------------------------------------------------------------------ tile finished: write it, take the next one
        if (__all_sync(FULL, phase == C_FETCH && s >= spp_eff)) {
            if (have_tile && active) {  // this work item's group sum; cloud_resolve_kernel adds the groups in order
                float *pp = P.partials + ((item_tile * 32 + lane) * FT_CLOUD_SAMPLE_GROUPS + item_group) * 3;
                pp[0] = sum.x;
                pp[1] = sum.y;
                pp[2] = sum.z;
            }
            // work item = (tile, sample group): the groups of one tile are consecutive items, so they run side by side on
            // different warps instead of one warp running all samples of its 32 pixels
            unsigned long long item = 0;
            if (lane == 0) item = atomicAdd(P.work_counter, 1ull);
            item = __shfl_sync(FULL, item, 0);
            if ((long long)item >= P.n_tiles * FT_CLOUD_SAMPLE_GROUPS) break;
            const unsigned long long tile = item / FT_CLOUD_SAMPLE_GROUPS;
            item_tile = (long long)tile;
            item_group = (int)(item - tile * FT_CLOUD_SAMPLE_GROUPS);
            active = cloud_tile_pixel(P, tiles_y, tile, lane, x, y);
            pixel = (uint32_t)((long long)x * P.height + y);
            s = (int)(((long long)item_group * P.spp) / FT_CLOUD_SAMPLE_GROUPS);
            spp_eff = active ? (int)(((long long)(item_group + 1) * P.spp) / FT_CLOUD_SAMPLE_GROUPS) : 0;
            if (!active) s = 0;
            sum = V3{0.f, 0.f, 0.f};
            have_tile = true;
            continue;
        }
        // ------------------------------------------------------------------ per-lane refill: the lane's next sample
        // (a finished lane waits until FT_CLOUD_REFILL_MIN lanes want their next sample, or nobody is busy: the set-up below --
        // Philox, three IEEE divisions of the normalisation -- ran with two lanes active on average when every lane started its
        // next sample the moment it finished the last, a tenth of the kernel's warp instructions)
        const bool want = phase == C_FETCH && s < spp_eff;
#if FT_CLOUD_REFILL_MIN > 1
        const unsigned wm = __ballot_sync(FULL, want), busy = __ballot_sync(FULL, phase != C_FETCH);
        if (want && (__popc(wm) >= FT_CLOUD_REFILL_MIN || busy == 0u)) {
#else
        if (want) {
#endif
            float u0, u1;
            cloud_rnd(P, spare, pixel, (uint32_t)s, 0u, u0, u1);  // sub-pixel jitter
            const float pu = ((float)x + u0) / P.max_size, pv = ((float)y + u1) / P.max_size;
            const V3 fwd = V3{P.cam.forward[0], P.cam.forward[1], P.cam.forward[2]};
            const V3 rs = V3{P.cam.right_scaled[0], P.cam.right_scaled[1], P.cam.right_scaled[2]};
            const V3 us = V3{P.cam.up_scaled[0], P.cam.up_scaled[1], P.cam.up_scaled[2]};
            d = normalize((fwd + rs * (pu - 0.5f)) + us * (pv - 0.5f));
            o = cam;
            len = P.length;
            steps = 0;
            if (len > 0.0f && d0 >= eps) {  // first iteration of the march with the shared Distance
                o = o + d * d0;
                len = len - d0;
                steps = 1;
            }
            seg = 0;
            phase = C_MARCH;
            add_flops<COUNT>(acc, FLOPS_SAMPLE_RAY);
        }
        // ------------------------------------------------------------------ pre-checks
        // event: 1 march miss, 2 march hit, 3 shadow unoccluded, 4 shadow occluded, 5 normal done, 6 sample ends on the
        // background, 7 light visible, 8 light hidden
        int event = 0;
        float t = 0.f;
        V3 p = o;
        if (phase == C_MARCH || phase == C_SHADOW) {
            if (len <= 0.0f || steps >= S.step_cap) {
                if (!(len <= 0.0f)) n_capped++;
                event = phase == C_MARCH ? 1 : 3;
            }
        } else if (phase >= C_N0 && phase <= C_N3) {
            const float h = eps * 0.125f;
            p = q;
            if (phase == C_N0) p.x = q.x + h;
            else if (phase == C_N1) p.y = q.y + h;
            else if (phase == C_N2) p.z = q.z + h;
        } else if (phase == C_TRACK) {
            t = ((float)k + xi0) * P.step;
            if (len - t <= 0.0f || k >= (uint32_t)FT_CLOUD_MAX_TRACK_STEPS) event = 6;
            p = o + d * t;
        } else if (phase == C_LTRACK) {
            t = ((float)k + xi0) * P.step;
            if (t >= maxdist) event = 7;
            else if (k >= (uint32_t)FT_CLOUD_MAX_TRACK_STEPS) event = 8;
            p = ps + d * t;
        }
        // chain scenes: one box test on the cell coordinate decides "no escape pre-check" and "straight to the sub-grid"
        const V3 cc = CHAIN ? chain_coord(S, p) : V3{0.f, 0.f, 0.f};
        const bool fast = CHAIN && in_fast_region(S, cc);
        if ((phase == C_MARCH || phase == C_SHADOW) && event == 0 && S.esc_valid && !fast && outside_box(S.esc_lo, S.esc_hi, o) &&
            ray_escapes(S, o, d, eps)) {
            n_escaped++;
            event = phase == C_MARCH ? 1 : 3;
        }
        // ------------------------------------------------------------------ the one distance evaluation
        if (phase > C_FETCH && event == 0) {
            uint32_t best_lo = 0u;
            const float dist = CHAIN ? eval_chain_c<COUNT>(S, p, cc, fast, acc, best_lo) : eval_form<COUNT>(S, S.root_word, p, acc);
            if (phase <= C_SHADOW) {
                steps++;
                add_flops<COUNT>(acc, FLOPS_MOVE + 1);
                if (dist < eps) {
                    event = phase == C_MARCH ? 2 : 4;
                    if (phase == C_MARCH) hit_best = best_lo;  // the hit evaluation's arg-min candidate
                } else {
                    o = o + d * dist;
                    len = len - dist;
                }
            } else if (phase == C_N0) {
                n.x = dist;
                phase = C_N1;
            } else if (phase == C_N1) {
                n.y = dist;
                phase = C_N2;
            } else if (phase == C_N2) {
                n.z = dist;
                phase = C_N3;
            } else if (phase == C_N3) {
                n = normalize(n - V3{dist, dist, dist});
                add_flops<COUNT>(acc, FLOPS_NORMAL);
                event = 5;
            } else {  // tracking step inside the medium
                n_track++;
                add_flops<COUNT>(acc, FLOPS_TRACK_STEP);
                if (dist >= eps) {
                    if (phase == C_TRACK) {  // left the medium: the camera ray goes on as an ordinary ray
                        seg++;
                        if (seg >= (uint32_t)FT_CLOUD_MAX_SEGMENTS) {
                            event = 6;
                        } else {
                            n_secondary++;
                            o = p;
                            len = len - t;
                            steps = 0;
                            phase = C_MARCH;
                        }
                    } else {  // the reference's shadow ray from here (SdfLight.fs:10-20 / 26-41)
                        n_secondary++;
                        const float4 lv = ld4(S.lights + 2 * light);
                        o = p;
                        if (__float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL) {
                            d = xyz(lv);
                            len = 1000.0f;
                        } else {
                            const V3 diff = xyz(lv) - p;
                            const float d2 = length2(diff);
                            d = diff / d2;
                            len = sqrtf(d2);
                        }
                        steps = 0;
                        phase = C_SHADOW;
                    }
                } else {
                    if (k > 0u) ucur = cloud_track_u(P, spare, pixel, (uint32_t)s, k, stream + (phase == C_LTRACK ? 1u + (uint32_t)light : 0u));
                    add_flops<COUNT>(acc, FLOPS_DENSITY);
                    if (ucur < cloud_density(tab, P, p, dist) * P.step) {
                        if (phase == C_TRACK) {  // real collision: single scattering at p
                            n_scattered++;
                            ps = p;
                            lc = bg;
                            light = 0;
                            cloud_mode = true;
                            event = 9;
                        } else {
                            event = 8;
                        }
                    } else {
                        k++;
                    }
                }
            }
        }
        // ------------------------------------------------------------------ transitions
        if (event != 0) {
            bool next_light = false, finish = false;
            V3 L = bg;
            if (event == 1 || event == 6) {  // SdfScene.fs:10
                finish = true;
            } else if (event == 2) {  // hit: the un-moved hit ray stays in (o, d, len), probe the normal
                q = o + d * (-eps);
                n_hits++;
                phase = C_N0;
            } else if (event == 5) {
                const int leaf_obj = (FT_CHAIN_MAT && CHAIN && S.chain_mat_valid) ? chain_leaf_from_best(S, o, hit_best)
                                                                                  : material_leaf<COUNT>(S, o, acc);
                mat = __ldg(S.objects + leaf_obj).z;
                if (__float_as_uint(ld4(S.materials + mat).w) == (uint32_t)FT_MATERIAL_CLOUD) {
                    stream = 1u + seg * nstreams;
                    cloud_rnd(P, spare, pixel, (uint32_t)s, stream, ucur, xi0);
                    k = 0u;
                    phase = C_TRACK;
                } else {  // the reference's shading, SdfScene.fs:12-28
                    lc = bg;
                    light = 0;
                    cloud_mode = false;
                    next_light = true;
                }
            } else if (event == 9) {
                next_light = true;
            } else {  // a light is decided: 3 / 7 it contributes, 4 / 8 it does not
                if (event == 3 || event == 7) {  // the light's intensity (and cosine) again from its record: same inputs, same operations
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    const bool directional = __float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL;
                    if (cloud_mode) {
                        lc = lc + (directional ? xyz(lrgb) : xyz(lrgb) / length2(xyz(lv) - ps));
                    } else {
                        const float lcos = dot(n, directional ? xyz(lv) : normalize(xyz(lv) - q));
                        lc = lc + (directional ? xyz(lrgb) : xyz(lrgb) / length2(xyz(lv) - q)) * lcos;
                    }
                }
                add_flops<COUNT>(acc, 6);
                light++;
                next_light = true;
            }
            if (next_light) {
                finish = true;
                for (; light < S.n_lights; light++) {
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    const bool directional = __float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL;
                    add_flops<COUNT>(acc, directional ? FLOPS_LIGHT : FLOPS_LIGHT + FLOPS_POINT_LIGHT);
                    if (cloud_mode) {  // isotropic scattering: every light, visibility by tracking
                        maxdist = __int_as_float(0x7f800000);
                        if (directional) {
                            d = xyz(lv);
                        } else {
                            const V3 diff = xyz(lv) - ps;
                            const float d2 = length2(diff);
                            maxdist = sqrtf(d2);
                            d = diff / maxdist;
                        }
                        n_secondary++;
                        cloud_rnd(P, spare, pixel, (uint32_t)s, stream + 1u + (uint32_t)light, ucur, xi0);
                        k = 0u;
                        phase = C_LTRACK;
                        finish = false;
                        break;
                    }
                    const V3 ldir = directional ? xyz(lv) : normalize(xyz(lv) - q);  // SdfScene.fs:13-19
                    const float lcos = dot(n, ldir);
                    if (lcos > 0.0f) {
                        o = q;
                        if (directional) {
                            d = xyz(lv);
                            len = 1000.0f;
                        } else {
                            const V3 diff = xyz(lv) - q;
                            const float d2 = length2(diff);
                            len = sqrtf(d2);
                            d = diff / d2;
                        }
                        steps = 0;
                        n_secondary++;
                        phase = C_SHADOW;
                        finish = false;
                        break;
                    }
                }
                if (finish) {
                    const V3 albedo = xyz(ld4(S.materials + mat));
                    const float pi_inv = 1.0f / 3.14159274f;
                    L = cloud_mode ? albedo * (lc * (pi_inv * 0.25f)) : albedo * (lc * pi_inv);
                    add_flops<COUNT>(acc, FLOPS_FINAL);
                }
            }
            if (finish) {
                sum = sum + L;
                add_flops<COUNT>(acc, 3);
                s++;
                phase = C_FETCH;
            }
        }
    }

    const unsigned hits = __reduce_add_sync(FULL, n_hits), secondary = __reduce_add_sync(FULL, n_secondary);
    const unsigned capped = __reduce_add_sync(FULL, n_capped), escaped = __reduce_add_sync(FULL, n_escaped);
    const unsigned scattered = __reduce_add_sync(FULL, n_scattered);
    for (int sh = 16; sh > 0; sh >>= 1) n_track += __shfl_xor_sync(FULL, n_track, sh);
    if (lane == 0) {
        const unsigned w = (blockIdx.x * (CB / 32) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1);
        unsigned long long *st = P.stats + w * FT_STAT_STRIDE;
        if (hits) atomicAdd(st + FT_STAT_HITS, (unsigned long long)hits);
        if (secondary) atomicAdd(st + FT_STAT_SHADOW, (unsigned long long)secondary);
        if (capped) atomicAdd(st + FT_STAT_CAPPED, (unsigned long long)capped);
        if (escaped) atomicAdd(st + FT_STAT_ESCAPED, (unsigned long long)escaped);
        if (scattered) atomicAdd(st + FT_STAT_PATHS, (unsigned long long)scattered);
        if (n_track) atomicAdd(st + FT_STAT_STEPS, n_track);
    }
    if (COUNT) {
        unsigned long long fl = acc.flops;
        for (int sh = 16; sh > 0; sh >>= 1) fl += __shfl_xor_sync(FULL, fl, sh);
        if (lane == 0) atomicAdd(P.stats + ((blockIdx.x * (CB / 32) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1)) * FT_STAT_STRIDE + FT_STAT_FLOPS, fl);
    }
}

}  // namespace ft_cloud

template <bool CHAIN, bool COUNT>
static cudaError_t cloud_launch_variant(const DevScene &S, const CloudParams &P, int sm_count, cudaStream_t stream) {
    using namespace ft_cloud;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cloud_kernel<CHAIN, COUNT>, CB, 0);
    if (e != cudaSuccess) return e;
    long long grid = (long long)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const long long need = (P.n_tiles * FT_CLOUD_SAMPLE_GROUPS + CB / 32 - 1) / (CB / 32);
    if (grid > need) grid = need;
    cloud_kernel<CHAIN, COUNT><<<(unsigned)grid, CB, 0, stream>>>(S, P);
    return cudaGetLastError();
}

// count: the instrumented variant (algorithmic flops into the stats block)
cudaError_t cloud_launch(const DevScene &S, const CloudParams &P, bool count, int sm_count, cudaStream_t stream) {
    using namespace ft_cloud;
    if (P.n_tiles <= 0) return cudaSuccess;
    cudaError_t e;
    if (S.chain_valid) e = count ? cloud_launch_variant<true, true>(S, P, sm_count, stream) : cloud_launch_variant<true, false>(S, P, sm_count, stream);
    else e = count ? cloud_launch_variant<false, true>(S, P, sm_count, stream) : cloud_launch_variant<false, false>(S, P, sm_count, stream);
    if (e != cudaSuccess) return e;
    cloud_resolve_kernel<<<(unsigned)((P.n_tiles * 32 + CB - 1) / CB), CB, 0, stream>>>(P);
    return cudaGetLastError();
}
