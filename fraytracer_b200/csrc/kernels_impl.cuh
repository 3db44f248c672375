// kernels_impl.cuh -- the sm_100a kernels of the hot path and their launchers.  Compiled twice:
// kernels_strict.cu (FT_NS = ft_strict, -fmad=false) and kernels_fast.cu (FT_NS = ft_fast, -fmad=true).
#pragma once
#include "kernels.h"
#include "trace_device.cuh"

namespace FT_NS {

constexpr int BLOCK = 128;  // 4 warps; each warp owns one 4 (x) by 8 (y) pixel tile

// stats slots (unsigned long long each), spread over STAT_SLOTS copies to keep the REDs off one address
__device__ __forceinline__ void flush_stats(unsigned long long *stats, bool valid, const TraceOut &t, const Acc &acc,
                                            bool count) {
    const unsigned hits = __reduce_add_sync(0xffffffffu, (valid && t.leaf >= 0) ? 1u : 0u);
    const unsigned shadows = __reduce_add_sync(0xffffffffu, valid ? t.n_shadow : 0u);
    const unsigned capped = __reduce_add_sync(0xffffffffu, valid ? t.capped : 0u);
    unsigned long long steps = valid ? acc.steps : 0ull, flops = valid ? acc.flops : 0ull;
    if (count) {
        for (int o = 16; o > 0; o >>= 1) {
            steps += __shfl_xor_sync(0xffffffffu, steps, o);
            flops += __shfl_xor_sync(0xffffffffu, flops, o);
        }
    }
    if ((threadIdx.x & 31) == 0) {
        const unsigned warp = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1);
        unsigned long long *s = stats + warp * FT_STAT_STRIDE;
        if (hits) atomicAdd(s + FT_STAT_HITS, (unsigned long long)hits);
        if (shadows) atomicAdd(s + FT_STAT_SHADOW, (unsigned long long)shadows);
        if (capped) atomicAdd(s + FT_STAT_CAPPED, (unsigned long long)capped);
        if (count) {
            atomicAdd(s + FT_STAT_STEPS, steps);
            atomicAdd(s + FT_STAT_FLOPS, flops);
        }
    }
}

// Image.render (Image.fs:26-35): one thread per pixel, primary-ray generation fused with the whole
// SdfScene.trace of that pixel.  Columns are addressed in the rank's compact strip space (see
// ft_render_device in the header); world == 1 degenerates to the plain x-major FColor[X,Y].
template <bool COUNT>
__global__ void __launch_bounds__(BLOCK) render_kernel(const __grid_constant__ DevScene S, const __grid_constant__ ft_camera cam,
                                                       RenderParams P, float *__restrict__ out_rgb,
                                                       int *__restrict__ out_leaf, unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int tiles_y = (P.height + 7) >> 3;
    const int tx = (int)(tile / tiles_y), ty = (int)(tile % tiles_y);
    const int xl = tx * 4 + (lane >> 3);  // column in the rank's compact space
    const int y = ty * 8 + (lane & 7);
    const int strip = xl / P.strip_width;
    const int x = P.x_offset + (strip * P.world + P.rank) * P.strip_width + (xl - strip * P.strip_width);  // image column
    const bool in_buffer = xl < P.local_cols && y < P.height;
    const bool valid = in_buffer && x < P.width;
    Acc acc{0ull, 0ull};
    TraceOut t{};
    t.leaf = -1;
    if (valid) {
        add_flops<COUNT>(acc, FLOPS_PRIMARY_RAY);
        const V3 dir = pixel_direction(cam, x, y, P.max_size);
        t = scene_trace<COUNT>(S, V3{cam.position[0], cam.position[1], cam.position[2]}, dir, P.length, P.epsilon, acc);
    }
    if (in_buffer) {
        const long long idx = (long long)xl * P.height + y;
        out_rgb[3 * idx + 0] = t.r;
        out_rgb[3 * idx + 1] = t.g;
        out_rgb[3 * idx + 2] = t.b;
        if (out_leaf) out_leaf[idx] = t.leaf;
    }
    flush_stats(stats, valid, t, acc, COUNT);
}

// batched SdfScene.trace for arbitrary rays: 32-byte ft_ray records, two LDG.128 per thread
template <bool COUNT>
__global__ void __launch_bounds__(BLOCK) trace_rays_kernel(const __grid_constant__ DevScene S, const float4 *__restrict__ rays,
                                                           long long n, float *__restrict__ out_rgb,
                                                           int *__restrict__ out_leaf, int *__restrict__ out_steps,
                                                           unsigned *__restrict__ out_shadow, unsigned long long *stats) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < n;
    Acc acc{0ull, 0ull};
    TraceOut t{};
    t.leaf = -1;
    if (valid) {
        const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
        t = scene_trace<COUNT>(S, V3{a.x, a.y, a.z}, V3{a.w, b.x, b.y}, b.z, b.w, acc);
        if (out_rgb) {
            out_rgb[3 * i + 0] = t.r;
            out_rgb[3 * i + 1] = t.g;
            out_rgb[3 * i + 2] = t.b;
        }
        if (out_leaf) out_leaf[i] = t.leaf;
        if (out_steps) out_steps[i] = t.steps;
        if (out_shadow) out_shadow[i] = t.shadow;
    }
    flush_stats(stats, valid, t, acc, COUNT);
}

// scene.Object.Form.Distance for n positions
__global__ void __launch_bounds__(BLOCK) distance_kernel(const __grid_constant__ DevScene S, const float *__restrict__ pts,
                                                         long long n, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    Acc acc{0ull, 0ull};
    out[i] = eval_form<false>(S, S.root_word, V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, acc);
}

cudaError_t launch_render(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *out_rgb, int *out_leaf,
                          unsigned long long *stats, bool count, cudaStream_t stream) {
    const long long tiles = (long long)((P.local_cols + 3) / 4) * ((P.height + 7) / 8);
    const long long blocks = (tiles + (BLOCK / 32) - 1) / (BLOCK / 32);
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (count)
        render_kernel<true><<<(unsigned)blocks, BLOCK, 0, stream>>>(S, cam, P, out_rgb, out_leaf, stats);
    else
        render_kernel<false><<<(unsigned)blocks, BLOCK, 0, stream>>>(S, cam, P, out_rgb, out_leaf, stats);
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(const DevScene &S, const float4 *rays, long long n, float *out_rgb, int *out_leaf,
                              int *out_steps, unsigned *out_shadow, unsigned long long *stats, bool count,
                              cudaStream_t stream) {
    const long long blocks = (n + BLOCK - 1) / BLOCK;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (count)
        trace_rays_kernel<true><<<(unsigned)blocks, BLOCK, 0, stream>>>(S, rays, n, out_rgb, out_leaf, out_steps, out_shadow, stats);
    else
        trace_rays_kernel<false><<<(unsigned)blocks, BLOCK, 0, stream>>>(S, rays, n, out_rgb, out_leaf, out_steps, out_shadow, stats);
    return cudaGetLastError();
}

cudaError_t launch_distance(const DevScene &S, const float *pts, long long n, float *out, cudaStream_t stream) {
    const long long blocks = (n + BLOCK - 1) / BLOCK;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    distance_kernel<<<(unsigned)blocks, BLOCK, 0, stream>>>(S, pts, n, out);
    return cudaGetLastError();
}

}  // namespace FT_NS
