// kernels_impl.cuh -- the sm_100a kernels of the hot path and their launchers.  Compiled twice:
// kernels_strict.cu (FT_NS = ft_strict, -fmad=false) and kernels_fast.cu (FT_NS = ft_fast, -fmad=true).
#pragma once
#include "kernels.h"
#include "trace_machine.cuh"

namespace FT_NS {

constexpr int BLOCK = 128;

// scene.Object.Form.Distance for n positions
__global__ void __launch_bounds__(BLOCK) distance_kernel(const __grid_constant__ DevScene S, const float *__restrict__ pts,
                                                         long long n, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    Acc acc{0ull, 0ull};
    const V3 p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    uint32_t best = 0u;
    // chain scenes go through the evaluator the render path uses (sub-grid included), so that tests see its values
    out[i] = S.chain_valid ? eval_chain<false>(S, p, acc, best) : eval_form<false>(S, S.root_word, p, acc);
}

// persistent grid: resident CTAs per SM (occupancy API) x SM count, capped by the amount of work
template <class K>
static cudaError_t persistent_grid(K kernel, long long items, int sm_count, unsigned *grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MB, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * sm_count;
    const long long need_blocks = (items + MB - 1) / MB;  // never more threads than work items
    if (g > need_blocks) g = need_blocks;
    if (g < 1) g = 1;
    *grid = (unsigned)g;
    return cudaSuccess;
}

#define FT_LAUNCH(C, E, H, JOB, ITEMS)                                                                               \
    do {                                                                                                              \
        if ((e = persistent_grid(machine_kernel<C, E, H, JOB>, ITEMS, sm_count, &grid)) != cudaSuccess) return e;     \
        machine_kernel<C, E, H, JOB><<<grid, MB, 0, stream>>>(S, job, work_counter, stats);                           \
    } while (0)

template <class JOB>
static cudaError_t launch_render_job(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *out_rgb, int *out_leaf,
                                     unsigned long long *work_counter, unsigned long long *stats, bool count, int sm_count,
                                     cudaStream_t stream) {
    JOB job;
    job.cam = cam;
    job.P = P;
    job.out_rgb = out_rgb;
    job.out_leaf = out_leaf;
    job.n_items = (long long)((P.local_cols + FT_TILE_W - 1) / FT_TILE_W) * ((P.height + FT_TILE_H - 1) / FT_TILE_H) * 32;
    if (job.n_items <= 0) return cudaSuccess;
    unsigned grid = 1;
    cudaError_t e;
    const bool chain = S.chain_valid != 0;
    if (count && chain) FT_LAUNCH(true, true, true, JOB, job.n_items);
    else if (count) FT_LAUNCH(true, true, false, JOB, job.n_items);
    else if (chain) FT_LAUNCH(false, true, true, JOB, job.n_items);
    else FT_LAUNCH(false, true, false, JOB, job.n_items);
    return cudaGetLastError();
}

// tile_store: the destination is mapped host memory (ft_render's zero-copy path): whole tiles leave as 128-bit stores
cudaError_t launch_render(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *out_rgb, int *out_leaf,
                          unsigned long long *work_counter, unsigned long long *stats, bool count, int sm_count,
                          cudaStream_t stream, bool tile_store) {
    if (tile_store) return launch_render_job<RenderTileJob>(S, cam, P, out_rgb, out_leaf, work_counter, stats, count, sm_count, stream);
    return launch_render_job<RenderJob>(S, cam, P, out_rgb, out_leaf, work_counter, stats, count, sm_count, stream);
}

cudaError_t launch_render_gather(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *const *dst, int n_dst,
                                 unsigned long long *work_counter, unsigned long long *stats, bool count, int sm_count,
                                 cudaStream_t stream) {
    if (n_dst < 1 || n_dst > FT_MAX_GATHER_DST) return cudaErrorInvalidValue;
    GatherJob job;
    job.cam = cam;
    job.P = P;
    for (int k = 0; k < FT_MAX_GATHER_DST; k++) job.dst[k] = k < n_dst ? dst[k] : nullptr;
    job.n_dst = n_dst;
    job.n_items = (long long)((P.local_cols + FT_TILE_W - 1) / FT_TILE_W) * ((P.height + FT_TILE_H - 1) / FT_TILE_H) * 32;
    if (job.n_items <= 0) return cudaSuccess;
    unsigned grid = 1;
    cudaError_t e;
    const bool chain = S.chain_valid != 0;
    if (count && chain) FT_LAUNCH(true, true, true, GatherJob, job.n_items);
    else if (count) FT_LAUNCH(true, true, false, GatherJob, job.n_items);
    else if (chain) FT_LAUNCH(false, true, true, GatherJob, job.n_items);
    else FT_LAUNCH(false, true, false, GatherJob, job.n_items);
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(const DevScene &S, const float4 *rays, long long n, float *out_rgb, int *out_leaf,
                              int *out_steps, unsigned *out_shadow, unsigned long long *work_counter,
                              unsigned long long *stats, bool count, int sm_count, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    RaysJob job{rays, n, out_rgb, out_leaf, out_steps, out_shadow};
    unsigned grid = 1;
    cudaError_t e;
    // per-ray step counts requested: every march is followed to its end (no escape proof)
    const bool escape = out_steps == nullptr, chain = S.chain_valid != 0;
    if (count) {  // the instrumented variant is not performance critical: generic interpreter only
        if (escape) FT_LAUNCH(true, true, false, RaysJob, n);
        else FT_LAUNCH(true, false, false, RaysJob, n);
    } else if (escape) {
        if (chain) FT_LAUNCH(false, true, true, RaysJob, n);
        else FT_LAUNCH(false, true, false, RaysJob, n);
    } else {
        if (chain) FT_LAUNCH(false, false, true, RaysJob, n);
        else FT_LAUNCH(false, false, false, RaysJob, n);
    }
    return cudaGetLastError();
}
#undef FT_LAUNCH

cudaError_t launch_distance(const DevScene &S, const float *pts, long long n, float *out, cudaStream_t stream) {
    const long long blocks = (n + BLOCK - 1) / BLOCK;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    distance_kernel<<<(unsigned)blocks, BLOCK, 0, stream>>>(S, pts, n, out);
    return cudaGetLastError();
}

}  // namespace FT_NS
