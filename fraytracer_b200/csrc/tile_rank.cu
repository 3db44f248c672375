// tile_rank.cu -- the order in which the next frame's tiles are drawn (ft_api.cu TileHint; scheduling only, pixels do not
// depend on it).
//
// A persistent grid hands out 4x8-pixel tiles from one counter; a tile is a serial chain of up to hundreds of Distance
// evaluations and a warp gets only about a dozen tiles per frame, so in natural order the launch ends when the last long tile
// does (ncu: sm__cycles_active.avg 82 % of elapsed on config 2).  Every launch records each tile's cost (the SM clock time
// its warp spent on it); this file turns the costs of one frame into the order of a later frame of the same shape:
//
//  * FT_RANK_LPT   -- device destinations: longest processing time first (a stable descending radix sort of the costs).
//  * FT_RANK_TIMED -- destinations behind PCIe (mapped host memory): the launch is then also bound by the rate at which
//    finished tiles leave over the link (24.9 MB per 1080p frame at <= 50 GB/s is 0.5 ms of a 0.8 ms launch), so the cheap
//    tiles must not all finish at the end.  Tiles dearer than `split` x the mean cost ("tops") keep their LPT order and get
//    the key  t = (cost of the tops before them) / (cost of all tops)  -- the moment they start if time ran on top cost alone;
//    the other tiles get keys spread uniformly over [0, 1) in natural (x-major) order, so that their stores flow at a constant
//    rate through the whole launch, vertical neighbours at about the same time (they fill whole lines on the way out:
//    tools/micro/pcie_store.cu: 4x8 tiles stored in scattered order reach 26-29 GB/s, in natural order 47 GB/s, the copy engine
//    56 GB/s).  One ascending radix sort of the keys gives the order.  Measured end to end through ft_render (B200, median of
//    40 frames, config 2 at 1080p / config 5 at 4K): natural order 1.11 / 2.99 ms, plain LPT 1.03 / 3.70, split 0.5 1.11 / 3.68,
//    split 1 0.97 / 3.22, split 1.5 0.99 / 3.05 (the default), split 2 1.01 / 3.02, split 4 1.05 / 3.11.  The launch itself
//    takes 0.914 ms with host stores in this order against 0.79 ms into device memory (LPT); ending the launch on its cheapest
//    tiles (a window for the non-tops that shrinks with their cost) changes nothing (0.913-0.921 ms): what is left is the link.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "kernels.h"

struct TileRank {
    unsigned n_tiles = 0, n_units = 0;
    int shift = 0;
    unsigned short *snap[2] = {nullptr, nullptr};  // unit costs of the frames of either parity
    unsigned *order[2] = {nullptr, nullptr};
    unsigned short *keys16 = nullptr;
    unsigned *iota = nullptr, *by_cost = nullptr, *flag = nullptr, *rank = nullptr, *keys32 = nullptr, *keys32_sorted = nullptr;
    unsigned long long *top_cost = nullptr, *top_prefix = nullptr, *sum = nullptr;
    void *temp = nullptr;
    size_t temp_bytes = 0;
};

namespace {

__global__ void __launch_bounds__(256) iota_kernel(unsigned *__restrict__ v, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// unit u = tiles [u << shift, (u + 1) << shift): saturating sum of their costs
__global__ void __launch_bounds__(256) unit_cost_kernel(const unsigned short *__restrict__ cost, unsigned n, int shift,
                                                        unsigned short *__restrict__ out, unsigned n_units) {
    const unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    unsigned sum = 0u;
    for (unsigned t = u << shift; t < ((u + 1u) << shift) && t < n; t++) sum += cost[t];
    out[u] = (unsigned short)(sum < 65535u ? sum : 65535u);
}

__global__ void __launch_bounds__(256) cost_sum_kernel(const unsigned short *__restrict__ c, unsigned n, unsigned long long *sum) {
    unsigned long long v = 0ull;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += c[i];
    for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(sum, v);
}

// a unit is a "top" when cost * n * 100 > sum * split_pct, i.e. cost > split x mean
__device__ __forceinline__ bool is_top(unsigned cost, unsigned n, unsigned long long sum, unsigned split_pct) {
    return (unsigned long long)cost * n * 100ull > sum * split_pct;
}

// by tile id: flag[u] = 1 for the non-tops (their natural-order rank is the exclusive scan of the flags)
__global__ void __launch_bounds__(256) flag_kernel(const unsigned short *__restrict__ c, unsigned n, const unsigned long long *sum,
                                                   unsigned split_pct, unsigned *__restrict__ flag) {
    const unsigned u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u <= n) flag[u] = (u < n && !is_top(c[u], n, *sum, split_pct)) ? 1u : 0u;
}
// by descending cost: the tops' costs (their start times are the exclusive scan)
__global__ void __launch_bounds__(256) top_cost_kernel(const unsigned short *__restrict__ sorted_cost, unsigned n,
                                                       const unsigned long long *sum, unsigned split_pct,
                                                       unsigned long long *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = (i < n && is_top(sorted_cost[i], n, *sum, split_pct)) ? (unsigned long long)sorted_cost[i] : 0ull;
}
__global__ void __launch_bounds__(256) timed_key_kernel(const unsigned short *__restrict__ sorted_cost, const unsigned *__restrict__ by_cost,
                                                        unsigned n, const unsigned long long *sum, unsigned split_pct,
                                                        const unsigned long long *__restrict__ top_prefix,
                                                        const unsigned *__restrict__ rank, unsigned *__restrict__ key) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double scale = 1073741824.0;  // keys in [0, 2^30]
    if (is_top(sorted_cost[i], n, *sum, split_pct)) {
        const unsigned long long all = top_prefix[n];
        key[i] = (unsigned)((double)top_prefix[i] / (double)(all ? all : 1ull) * scale);
    } else {
        const unsigned n_rest = rank[n];
        key[i] = (unsigned)(((double)rank[by_cost[i]] + 0.5) / (double)(n_rest ? n_rest : 1u) * scale);
    }
}

}  // namespace

void tile_rank_destroy(TileRank *r) {
    if (!r) return;
    void *ptrs[] = {r->snap[0], r->snap[1], r->order[0], r->order[1], r->keys16, r->iota, r->by_cost, r->flag, r->rank,
                    r->keys32, r->keys32_sorted, r->top_cost, r->top_prefix, r->sum, r->temp};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete r;
}

cudaError_t tile_rank_create(unsigned n_tiles, int shift, cudaStream_t stream, TileRank **out) {
    *out = nullptr;
    TileRank *r = new (std::nothrow) TileRank;
    if (!r) return cudaErrorMemoryAllocation;
    r->n_tiles = n_tiles;
    r->shift = shift;
    const unsigned n = r->n_units = (n_tiles + (1u << shift) - 1u) >> shift;
    size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, t1, (const unsigned short *)nullptr, (unsigned short *)nullptr,
                                              (const unsigned *)nullptr, (unsigned *)nullptr, (int)n, 0, 16);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, (int)n, 0, 31);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, (const unsigned *)nullptr, (unsigned *)nullptr, (int)(n + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, t4, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)(n + 1));
    r->temp_bytes = std::max(std::max(t1, t2), std::max(t3, t4)) + 256;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](auto **p, size_t count) {
        if (e == cudaSuccess) e = cudaMalloc((void **)p, sizeof(**p) * count);
    };
    for (int k = 0; k < 2; k++) {
        alloc(&r->snap[k], n);
        alloc(&r->order[k], n);
    }
    alloc(&r->keys16, n);
    alloc(&r->iota, n);
    alloc(&r->by_cost, n);
    alloc(&r->flag, (size_t)n + 1);
    alloc(&r->rank, (size_t)n + 1);
    alloc(&r->keys32, n);
    alloc(&r->keys32_sorted, n);
    alloc(&r->top_cost, (size_t)n + 1);
    alloc(&r->top_prefix, (size_t)n + 1);
    alloc(&r->sum, 1);
    if (e == cudaSuccess) e = cudaMalloc(&r->temp, r->temp_bytes);
    if (e == cudaSuccess) {
        iota_kernel<<<(n + 255) / 256, 256, 0, stream>>>(r->iota, n);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        tile_rank_destroy(r);
        return e;
    }
    *out = r;
    return cudaSuccess;
}

const unsigned *tile_rank_order(const TileRank *r, int parity) { return r->order[parity & 1]; }
int tile_rank_shift(const TileRank *r) { return r->shift; }

cudaError_t tile_rank_snapshot(TileRank *r, const unsigned short *cost, int parity, cudaStream_t stream) {
    unit_cost_kernel<<<(r->n_units + 255) / 256, 256, 0, stream>>>(cost, r->n_tiles, r->shift, r->snap[parity & 1], r->n_units);
    return cudaGetLastError();
}

cudaError_t tile_rank_run(TileRank *r, int parity, int mode, int split_pct, cudaStream_t stream) {
    const unsigned n = r->n_units;
    const unsigned short *snap = r->snap[parity & 1];
    unsigned *order = r->order[parity & 1];
    const unsigned blocks = (n + 256) / 256;  // covers n + 1 elements
    cudaError_t e;
    if (mode != FT_RANK_TIMED)
        return cub::DeviceRadixSort::SortPairsDescending(r->temp, r->temp_bytes, snap, r->keys16, r->iota, order, (int)n, 0, 16, stream);
    if ((e = cub::DeviceRadixSort::SortPairsDescending(r->temp, r->temp_bytes, snap, r->keys16, r->iota, r->by_cost, (int)n, 0, 16,
                                                      stream)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(r->sum, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    cost_sum_kernel<<<std::min(blocks, 592u), 256, 0, stream>>>(snap, n, r->sum);
    flag_kernel<<<blocks, 256, 0, stream>>>(snap, n, r->sum, (unsigned)split_pct, r->flag);
    top_cost_kernel<<<blocks, 256, 0, stream>>>(r->keys16, n, r->sum, (unsigned)split_pct, r->top_cost);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(r->temp, r->temp_bytes, r->flag, r->rank, (int)(n + 1), stream)) != cudaSuccess) return e;
    if ((e = cub::DeviceScan::ExclusiveSum(r->temp, r->temp_bytes, r->top_cost, r->top_prefix, (int)(n + 1), stream)) != cudaSuccess)
        return e;
    timed_key_kernel<<<blocks, 256, 0, stream>>>(r->keys16, r->by_cost, n, r->sum, (unsigned)split_pct, r->top_prefix, r->rank, r->keys32);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(r->temp, r->temp_bytes, r->keys32, r->keys32_sorted, r->by_cost, order, (int)n, 0, 31, stream);
}
