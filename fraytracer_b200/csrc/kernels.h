// kernels.h -- launcher declarations shared by ft_api.cu and the two kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/fraytracer_b200.h"
#include "device_scene.h"

#define FT_STAT_SLOTS 64   // power of two
#define FT_STAT_STRIDE 16  // unsigned long long per slot (128 B apart)
#define FT_STAT_HITS 0
#define FT_STAT_SHADOW 1
#define FT_STAT_CAPPED 2
#define FT_STAT_STEPS 3
#define FT_STAT_FLOPS 4
#define FT_STAT_ESCAPED 5
#define FT_STAT_LONG_FLOPS 7  // COUNT variant: work of the evaluations outside the largest grid's box
#define FT_STAT_LONG_EVALS 8
#define FT_STAT_EVALS 9
#define FT_STAT_PATHS 6  // spectral extension: continuation path segments marched

struct RenderParams {
    int width, height;   // image X, Y
    int local_cols;      // columns of this rank's compact buffer
    int strip_width, rank, world;
    int x_offset;        // image column of compact column 0 (band rendering)
    int xl_begin;        // first compact column (in the rank's strip space) this launch renders; out_rgb starts there
    float epsilon, length, max_size;
    // Scheduling hints (ft_api.cu TileHint; never used in arithmetic): `order` = the launch's tiles, most expensive first, as
    // measured by the previous launch of the same shape on this scene (null: natural order); `cost` receives the time
    // the warp spent on every tile (SM clock, units of 64 cycles, saturating) for the later launches (null: not recorded).
    const unsigned *order;
    unsigned short *cost;
    int order_mode;      // 0: order[] as it is (most expensive first); 1: expensive and cheap units alternate (order[0], order[n-1],
                         // order[1], order[n-2] ...) -- for destinations behind PCIe, where the cheap tiles' stores must not all
                         // arrive at the end of the launch
    int order_shift;     // order[] ranks units of (1 << order_shift) consecutive tiles (vertical neighbours: tile ids run along y).
                         // 0 in every shipped path.  (Units were meant for destinations behind PCIe, where neighbouring 96-byte
                         // column runs leave the GPU as full lines when they are written at about the same time -- tools/micro/
                         // pcie_store.cu: 4x8 tiles in scattered order 29 GB/s, in natural order 47 GB/s, 4x32 units in scattered
                         // order 46-48 GB/s -- but in ft_render the timed merge of single tiles measured best, ft_api.cu.)
};

// Which pixels a rank renders and where they go (multi-GPU): column strips dealt round-robin
// (strip s -> rank s % world) and a list of destination framebuffers -- the rank's own, its peers' (CUDA IPC) or mapped
// host memory -- that receive every finished pixel at its final x-major position.  world == 1: the whole frame.
#define FT_MAX_DST 16
struct StripTarget {
    int strip_width, rank, world;  // strip_width == 0: 4 x 8 pixel tiles dealt round-robin instead (tile t -> rank t % world)
    int local_cols;  // columns of the rank's compact strip space (strips_per_rank * strip_width; width when world == 1)
    int n_dst;
    float *dst[FT_MAX_DST];
#if defined(__CUDACC__)
    // column of the compact strip space -> image column (may be >= width: padding of the last strip)
    __device__ __forceinline__ int image_column(int xl) const {
        const int strip = xl / strip_width;
        return (strip * world + rank) * strip_width + (xl - strip * strip_width);
    }
    __device__ __forceinline__ void store(unsigned long long pixel, float r, float g, float b) const {
        for (int k = 0; k < n_dst; k++) {
            float *out = dst[k] + 3ull * pixel;
            out[0] = r;
            out[1] = g;
            out[2] = b;
        }
    }
#endif
};

#define FT_DECLARE_LAUNCHERS(NS)                                                                                      \
    namespace NS {                                                                                                    \
    cudaError_t launch_render(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *out_rgb,         \
                              int *out_leaf, unsigned long long *work_counter, unsigned long long *stats, bool count, \
                              int sm_count, cudaStream_t stream, bool tile_store);                                    \
    cudaError_t launch_render_gather(const DevScene &S, const ft_camera &cam, const RenderParams &P, float *const *dst,     \
                                     int n_dst, unsigned long long *work_counter, unsigned long long *stats, bool count,   \
                                     int sm_count, cudaStream_t stream);                                              \
    cudaError_t launch_trace_rays(const DevScene &S, const float4 *rays, long long n, float *out_rgb, int *out_leaf,  \
                                  int *out_steps, unsigned *out_shadow, unsigned long long *work_counter,             \
                                  unsigned long long *stats, bool count, int sm_count, cudaStream_t stream);          \
    cudaError_t launch_distance(const DevScene &S, const float *pts, long long n, float *out, cudaStream_t stream);   \
    }

FT_DECLARE_LAUNCHERS(ft_strict)
FT_DECLARE_LAUNCHERS(ft_strict_cone)
FT_DECLARE_LAUNCHERS(ft_fast)

// util_kernels.cu
cudaError_t launch_deinterleave(const float *gathered, int width, int height, int strip_width, int world, int channels,
                                int first_local_strip, int n_local_strips, float *out, cudaStream_t stream);
cudaError_t measure_fp32_peak(int device, float *tflops_fma, float *sm_mhz_estimate);

// tile_rank.cu: the order in which the next frames' tiles are drawn, from the per-tile costs a frame recorded
struct TileRank;
enum { FT_RANK_LPT = 0, FT_RANK_TIMED = 2 };
cudaError_t tile_rank_create(unsigned n_tiles, int shift, cudaStream_t stream, TileRank **out);
void tile_rank_destroy(TileRank *r);
// on the render stream, right after the kernel that wrote cost[]: keep the frame's unit costs (frames alternate in parity)
cudaError_t tile_rank_snapshot(TileRank *r, const unsigned short *cost, int parity, cudaStream_t stream);
// on any stream ordered after the snapshot: order[parity] = the units in drawing order
cudaError_t tile_rank_run(TileRank *r, int parity, int mode, int split_pct, cudaStream_t stream);
const unsigned *tile_rank_order(const TileRank *r, int parity);
int tile_rank_shift(const TileRank *r);

// post_kernels.cu
// strip_width > 0: only the columns of the strips `rank` owns (strip s -> rank s % world); 0: the whole frame
cudaError_t launch_image_max(const float *rgb, long long n_floats, float *d_max, cudaStream_t stream, int column_floats = 0,
                             int strip_width = 0, int rank = 0, int world = 1);
cudaError_t launch_tonemap(const float *rgb, int X, int Y, float gamma, const float *d_max, float max_value,
                           unsigned long long seed, unsigned char *out, int stride, cudaStream_t stream, int strip_width = 0,
                           int rank = 0, int world = 1);
int write_bmp_file(const char *path, const unsigned char *bgr, int width, int height, int stride);
