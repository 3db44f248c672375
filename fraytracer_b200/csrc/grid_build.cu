// grid_build.cu -- SdfBoundary.buildSpatialLookup (SdfBoundary.fs:225-274) on the device (SURVEY.md section 8(f)
// rank 2).  One thread per grid cell runs exactly the host compiler's float sequence (this file is compiled with
// -fmad=false, IEEE sqrt): cell centre (:246), upperBound = min_i getMaxDistance + |cellSize/2| (:248-253), the
// filter getMinDistance < upperBound (:255-264) and the .NET introsort by LowerBound (:267-268, dotnet_sort.h) --
// so the lists are bit-identical to the host build and to the oracle's (tests compare them).
// The bounds of all children are read by every thread at the same address (broadcast, L1 resident).
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "dotnet_sort.h"
#include "ft_math.h"
#include "scene_compiler.h"

namespace {
using namespace ftm;
using ftsort::KeyItem;
static_assert(sizeof(KeyItem) == sizeof(ftc::GridKeyItem), "KeyItem layout");

struct GridDev {
    float amin[3], cell[3], half[3], half_len;
    int cx, cy, cz, n;
};

__device__ __forceinline__ V3 cell_center(const GridDev &P, int c) {
    const int z = c % P.cz, y = (c / P.cz) % P.cy, x = c / (P.cz * P.cy);
    const V3 amin = v3(P.amin[0], P.amin[1], P.amin[2]), half = v3(P.half[0], P.half[1], P.half[2]);
    const V3 cell = v3(P.cell[0], P.cell[1], P.cell[2]);
    return (amin + half) + cell * v3((float)x, (float)y, (float)z);  // SdfBoundary.fs:246
}

__global__ void __launch_bounds__(128) grid_count_kernel(const float4 *__restrict__ b, GridDev P, int n_cells,
                                                         float *__restrict__ upper, int *__restrict__ count,
                                                         float *__restrict__ centers) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const V3 center = cell_center(P, c);
    float4 bb = __ldg(b);
    float m = distance(v3(bb.x, bb.y, bb.z), center) + bb.w;  // getMaxDistance, Seq.min (:248-252)
    for (int i = 1; i < P.n; i++) {
        bb = __ldg(b + i);
        const float dd = distance(v3(bb.x, bb.y, bb.z), center) + bb.w;
        if (dd < m) m = dd;
    }
    const float up = m + P.half_len;  // :253
    int k = 0;
    for (int i = 0; i < P.n; i++) {
        bb = __ldg(b + i);
        k += (distance(v3(bb.x, bb.y, bb.z), center) - bb.w) < up;  // getMinDistance < upperBound (:258)
    }
    upper[c] = up;
    count[c] = k;
    centers[3 * c] = center.x;
    centers[3 * c + 1] = center.y;
    centers[3 * c + 2] = center.z;
}

__global__ void __launch_bounds__(128) grid_fill_kernel(const float4 *__restrict__ b, GridDev P, int n_cells,
                                                        const float *__restrict__ upper, const long long *__restrict__ begin,
                                                        KeyItem *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const V3 center = cell_center(P, c);
    const float up = upper[c];
    KeyItem *list = out + begin[c];
    int k = 0;
    for (int i = 0; i < P.n; i++) {
        const float4 bb = __ldg(b + i);
        const float md = distance(v3(bb.x, bb.y, bb.z), center) - bb.w;
        if (md < up) list[k++] = KeyItem{md, i};
    }
    ftsort::sort(list, k);  // :267-268
}

#define GB_CU(call)                                                                          \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            if (err) *err = std::string("gpu grid build: ") + #call + ": " + cudaGetErrorString(e__); \
            cleanup();                                                                       \
            return FT_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

}  // namespace

namespace ftc {

int gpu_grid_build(void *ctx, const GridBuildJob &job, GridBuildResult *res, std::string *err) {
    const int device = (int)(intptr_t)ctx;
    int prev = -1;
    cudaGetDevice(&prev);
    float4 *d_b = nullptr;
    float *d_upper = nullptr, *d_centers = nullptr;
    int *d_count = nullptr;
    long long *d_begin = nullptr;
    KeyItem *d_out = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_b); cudaFree(d_upper); cudaFree(d_centers); cudaFree(d_count); cudaFree(d_begin); cudaFree(d_out);
        if (prev >= 0) cudaSetDevice(prev);
    };
    GB_CU(cudaSetDevice(device));
    const int n_cells = job.cx * job.cy * job.cz;
    GridDev P;
    for (int k = 0; k < 3; k++) { P.amin[k] = job.amin[k]; P.cell[k] = job.cell[k]; P.half[k] = job.half[k]; }
    P.half_len = job.half_len; P.cx = job.cx; P.cy = job.cy; P.cz = job.cz; P.n = job.n;
    GB_CU(cudaMalloc(&d_b, sizeof(float4) * (size_t)job.n));
    GB_CU(cudaMemcpy(d_b, job.bounds, sizeof(float4) * (size_t)job.n, cudaMemcpyHostToDevice));
    GB_CU(cudaMalloc(&d_upper, sizeof(float) * (size_t)n_cells));
    GB_CU(cudaMalloc(&d_centers, sizeof(float) * 3 * (size_t)n_cells));
    GB_CU(cudaMalloc(&d_count, sizeof(int) * (size_t)n_cells));
    GB_CU(cudaMalloc(&d_begin, sizeof(long long) * ((size_t)n_cells + 1)));
    const unsigned blocks = (unsigned)((n_cells + 127) / 128);
    grid_count_kernel<<<blocks, 128>>>(d_b, P, n_cells, d_upper, d_count, d_centers);
    GB_CU(cudaGetLastError());
    std::vector<int> count((size_t)n_cells);
    GB_CU(cudaMemcpy(count.data(), d_count, sizeof(int) * (size_t)n_cells, cudaMemcpyDeviceToHost));
    res->begin.assign((size_t)n_cells + 1, 0);
    for (int c = 0; c < n_cells; c++) res->begin[(size_t)c + 1] = res->begin[(size_t)c] + count[(size_t)c];
    const long long total = res->begin[(size_t)n_cells];
    GB_CU(cudaMemcpy(d_begin, res->begin.data(), sizeof(long long) * ((size_t)n_cells + 1), cudaMemcpyHostToDevice));
    GB_CU(cudaMalloc(&d_out, sizeof(KeyItem) * (size_t)(total > 0 ? total : 1)));
    grid_fill_kernel<<<blocks, 128>>>(d_b, P, n_cells, d_upper, d_begin, d_out);
    GB_CU(cudaGetLastError());
    res->entries.resize((size_t)total);
    res->centers.resize(3 * (size_t)n_cells);
    if (total > 0) GB_CU(cudaMemcpy(res->entries.data(), d_out, sizeof(KeyItem) * (size_t)total, cudaMemcpyDeviceToHost));
    GB_CU(cudaMemcpy(res->centers.data(), d_centers, sizeof(float) * 3 * (size_t)n_cells, cudaMemcpyDeviceToHost));
    cleanup();
    return FT_OK;
}

}  // namespace ftc
