// trace_machine.cuh -- the persistent "ray machine" kernel.
//
// Every lane runs a small state machine over ONE inlined call site of the distance evaluator:
//
//     FETCH -> primary march -> (hit) 4 normal probes -> material arg-min -> per light: shadow march -> store
//
// so that lanes of one warp that are in different phases of SdfScene.trace (one still marching its primary
// ray, one probing its normal, one following a shadow ray) still execute the same instructions.  A warp works on
// one chunk of FT_CHUNK items -- a 4 x 8 pixel tile -- at a time, all lanes in lockstep, and takes the next chunk from
// one global counter (one atomic per chunk) when every lane is idle: what an SM pays for is warp instructions and an
// iteration costs the maximum over its active lanes, so adjacent pixels (same cells, same candidate lists, similar
// trip counts) are worth more than busy lanes (per-lane refill was built first and measured 1.5-2.9x slower).
// The grid is persistent: resident-CTAs-per-SM x SM count, sized by the occupancy API.
#pragma once
#include "trace_device.cuh"

namespace FT_NS {

#ifndef FT_MINB
#define FT_MINB 9  // <= 56 registers: 36 resident warps per SM.  Measured (config 2): 6 CTAs/SM (80 registers, no spill) 2.25 ms,
                   // 8 (64) 1.88 ms, 9 (56) 1.83 ms
#endif
constexpr int MB = 128;             // threads per CTA
#ifndef FT_TILE_W
#define FT_TILE_W 4  // pixel tile of one warp: FT_TILE_W (x) by FT_TILE_H (y), 32 pixels
#define FT_TILE_H 8
#endif
#ifndef FT_LONG_VOTE
#define FT_LONG_VOTE 0  // 1: the warp runs evaluations outside / inside the largest grid's box in separate iterations.  It paid
                       // before the escape proof and whole-tile lockstep (outside lookups were 27 % of all); now they are 2.6 %
                       // and the two ballots per iteration cost more than they save: off is 2.5-4 % faster (measured)
#endif
#ifndef FT_LONG_BATCH
#define FT_LONG_BATCH 1  // run the long class as soon as any lane has one (measured best with whole-tile lockstep)
#endif
#ifndef FT_TILE_STORE
#define FT_TILE_STORE 1  // whole tiles leave as 128-bit stores (RenderJob / GatherJob::store_tile)
#endif
#ifndef FT_REFILL_MIN
#define FT_REFILL_MIN 32  // idle lanes that trigger a refill (see the refill block of machine_kernel)
#endif
constexpr int FT_CHUNK = 32;        // work items per global-counter grab (one tile of 4x8 pixels)
#ifndef FT_LAZY_WSTATS
#define FT_LAZY_WSTATS 1  // the address of the warp's event counters is formed inside the rare branch that bumps one (volatile asm
                          // reading %tid.x cannot be hoisted): ptxas otherwise rematerialises it -- S2R, S2R, SHF, LEA, LOP3, IADD3 --
                          // at the head of the march pre-check of EVERY iteration (1.6 % of the warp instructions, ncu r03z)
#endif

// one of the warp's four event counters (hits, shadow rays, capped rays, rays ended by the escape proof) in shared memory
__device__ __forceinline__ void warp_event(unsigned *wstats, unsigned sh_base, int k) {
#if FT_LAZY_WSTATS
    (void)wstats;
    asm volatile(
        "{\n\t"
        ".reg .u32 t;\n\t"
        "mov.u32 t, %%tid.x;\n\t"
        "shr.u32 t, t, 5;\n\t"
        "shl.b32 t, t, 4;\n\t"
        "add.u32 t, t, %0;\n\t"
        "red.shared.add.u32 [t], 1;\n\t"
        "}" ::"r"(sh_base + 4u * (unsigned)k)
        : "memory");
#else
    (void)sh_base;
    atomicAdd(wstats + k, 1u);
#endif
}

enum : int {
    PH_DONE = -1,
    PH_FETCH = 0,
    PH_PRIMARY = 1,   // marching the primary ray
    PH_SHADOW = 2,    // marching a shadow ray
    PH_N0 = 3,        // normal probes: +x, +y, +z, centre (SdfForm.fs:106-112)
    PH_N1 = 4,
    PH_N2 = 5,
    PH_N3 = 6,
    PH_PROBE = 7,     // Distance at the common origin of all primary rays (render jobs), once per warp
};

// Scheduling hint (RenderParams::order): the n-th chunk drawn from the work counter belongs to unit order[n >> shift] of the
// previous launch's cost ranking (a unit = 1 << shift vertically adjacent tiles), or -- order_mode 1 -- of the ranking read from
// both ends alternately.  The last unit may be partial: its missing tiles map beyond n_items and are skipped by the caller.
__device__ __forceinline__ long long hinted_chunk(const RenderParams &P, long long base, long long n_items) {
    if (!P.order) return base;
    const long long t = base >> 5;
    const int sh = P.order_shift;
    long long u = t >> sh;
    if (P.order_mode == 1) {
        const long long n_units = ((n_items >> 5) + (1ll << sh) - 1) >> sh;
        u = (u & 1) ? n_units - 1 - (u >> 1) : (u >> 1);
    }
    return (((long long)__ldg(P.order + u) << sh) + (t & ((1ll << sh) - 1))) << 5;
}
// chunks to draw: whole units when a ranking is attached
__device__ __forceinline__ long long hinted_total(const RenderParams &P, long long n_items) {
    if (!P.order) return n_items;
    const int sh = P.order_shift;
    return ((((n_items >> 5) + (1ll << sh) - 1) >> sh) << sh) << 5;
}

// ---- work sources -----------------------------------------------------------------------------------------
// Image.render: item i -> pixel, tiles of 4 (x) by 8 (y) pixels, tiles ordered along y first so that a chunk
// of 8 tiles is a 4 x 64 pixel block of the x-major FColor[X,Y] output.
template <bool TILE>
struct RenderJobT {
    ft_camera cam;
    RenderParams P;
    float *out_rgb;
    int *out_leaf;
    long long n_items;  // tiles * 32

    typedef int slot_t;  // pixels of one frame: < 2^31 (checked by the API)
    __device__ __forceinline__ long long total() const { return n_items; }
    // all primary rays start at the camera position: Distance there is evaluated once per warp
    static constexpr bool kProbe = true;
    static constexpr bool kUniformEps = true;  // every ray of the job has the same Epsilon: read it from the constant bank
    __device__ __forceinline__ float uniform_eps() const { return P.epsilon; }
    __device__ __forceinline__ V3 probe_point() const { return V3{cam.position[0], cam.position[1], cam.position[2]}; }
    // returns false when the item is padding (outside the image)
    __device__ __forceinline__ bool load(long long i, float d0, V3 &o, V3 &d, float &len, float &eps, slot_t &slot,
                                         int &steps) const {
        const long long tile = i >> 5;
        const int lane = (int)(i & 31);
        const int tiles_y = (P.height + FT_TILE_H - 1) / FT_TILE_H;
        const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
        const int xc = tx * FT_TILE_W + lane / FT_TILE_H;  // column inside this launch's buffer
        const int xl = P.xl_begin + xc;                    // column in the rank's compact strip space
        const int y = ty * FT_TILE_H + lane % FT_TILE_H;
        const int strip = xl / P.strip_width;
        const int x = P.x_offset + (strip * P.world + P.rank) * P.strip_width + (xl - strip * P.strip_width);
        if (xc >= P.local_cols || y >= P.height) return false;
        slot = xc * P.height + y;
        if (x >= P.width) {  // padding column of a strip buffer: defined contents for the all-gather
            out_rgb[3ll * slot + 0] = 0.f;
            out_rgb[3ll * slot + 1] = 0.f;
            out_rgb[3ll * slot + 2] = 0.f;
            if (out_leaf) out_leaf[slot] = -1;
            return false;
        }
        o = V3{cam.position[0], cam.position[1], cam.position[2]};
        d = pixel_direction(cam, x, y, P.max_size);
        len = P.length;
        eps = P.epsilon;
        steps = 0;
        // first iteration of SdfForm.tryTrace (SdfForm.fs:94-104) with the shared Distance d0 = D(camera):
        // Length > 0 and not a hit  =>  Ray.move d0.  Any other case is left to the generic path.
        if (len > 0.0f && d0 >= eps) {
            o = o + d * d0;
            len = len - d0;
            steps = 1;
        }
        return true;
    }
    __device__ __forceinline__ void store(slot_t slot, float r, float g, float b, int leaf, int, unsigned) const {
        out_rgb[3ll * slot + 0] = r;
        out_rgb[3ll * slot + 1] = g;
        out_rgb[3ll * slot + 2] = b;
        if (out_leaf) out_leaf[slot] = leaf;
    }
    // Whole-tile store (FT_TILE_STORE): `sh` holds the tile's 32 pixels in tile order (lane = column * 8 + row, 3 floats each),
    // i.e. the four 96-byte column runs of the x-major FColor[X,Y] back to back; `valid` marks the lanes that hold a pixel.
    // A complete, 16-byte-aligned tile leaves as 24 128-bit stores -- full 32-byte sectors for mapped host memory and peer
    // framebuffers (PCIe / NVLink write efficiency) -- anything else pixel by pixel.
    static constexpr bool kTileCost = true;
    __device__ __forceinline__ long long map_chunk(long long base) const { return hinted_chunk(P, base, n_items); }
    __device__ __forceinline__ long long draw_total() const { return hinted_total(P, n_items); }
    __device__ __forceinline__ void note_cost(long long base, unsigned cost) const {
        if (P.cost && base < n_items) P.cost[base >> 5] = (unsigned short)(cost < 65535u ? cost : 65535u);
    }
    static constexpr bool kTileStore = TILE;  // launch_render picks it for destinations behind PCIe (mapped host memory)
    __device__ __forceinline__ int *out_leaf_ptr() const { return out_leaf; }
    __device__ __forceinline__ void store_tile(long long base_item, const float *sh, unsigned valid, int lane) const {
        const long long tile = base_item >> 5;
        const int tiles_y = (P.height + FT_TILE_H - 1) / FT_TILE_H;
        const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
        const int xc0 = tx * FT_TILE_W, y0 = ty * FT_TILE_H;
        if (valid == 0xffffffffu && (P.height & 3) == 0 && ((uintptr_t)out_rgb & 15) == 0) {
            if (lane < 24) {
                const int col = lane / 6, part = lane - col * 6;
                const float4 v = *reinterpret_cast<const float4 *>(sh + col * 24 + part * 4);
                *reinterpret_cast<float4 *>(out_rgb + 3ll * ((long long)(xc0 + col) * P.height + y0) + part * 4) = v;
            }
        } else if ((valid >> lane) & 1u) {
            const long long slot = (long long)(xc0 + lane / FT_TILE_H) * P.height + (y0 + lane % FT_TILE_H);
            out_rgb[3 * slot + 0] = sh[3 * lane + 0];
            out_rgb[3 * slot + 1] = sh[3 * lane + 1];
            out_rgb[3 * slot + 2] = sh[3 * lane + 2];
        }
    }
};

typedef RenderJobT<false> RenderJob;      // device-memory destination: plain per-pixel stores (2-3 % faster on the device)
typedef RenderJobT<true> RenderTileJob;   // mapped host memory: whole tiles as 128-bit stores

// (Measured and rejected, round 2: staging finished tiles in an L2-resident device frame and letting the warp that completes the
// last tile of a group of four vertically adjacent tiles copy the group's 384-byte column runs to the host frame in one go --
// one atomic per tile, no waiting.  The PCIe microbenchmark favours long runs (96-byte runs in scattered order 26-29 GB/s, 384-byte
// runs 48-50 GB/s), yet ft_render got slower, config 2 0.870 -> 1.000 ms, config 5 at 4K 2.99 -> 3.08 ms: a tile's pixels then leave
// only when the slowest of four tiles is done, and the fence per tile waits behind the warp's own host stores.)

// Image.render of this rank's column strips, fused with the framebuffer gather: every finished
// pixel is stored at its final x-major position in the framebuffer of EVERY destination -- the rank's own and its
// peers', mapped through CUDA IPC, so the stores travel over NVLink / NVSwitch while the rest of the tile is still
// being traced.  No compact strip buffer, no all-gather, no de-interleave pass; a 4-byte all-reduce afterwards tells
// every rank that all writers have finished.
#define FT_MAX_GATHER_DST 16
struct GatherJob {
    ft_camera cam;
    RenderParams P;
    float *dst[FT_MAX_GATHER_DST];
    int n_dst;
    long long n_items;

    typedef int slot_t;
    __device__ __forceinline__ long long total() const { return n_items; }
    static constexpr bool kProbe = true;
    static constexpr bool kUniformEps = true;
    __device__ __forceinline__ float uniform_eps() const { return P.epsilon; }
    __device__ __forceinline__ V3 probe_point() const { return V3{cam.position[0], cam.position[1], cam.position[2]}; }
    __device__ __forceinline__ bool load(long long i, float d0, V3 &o, V3 &d, float &len, float &eps, slot_t &slot,
                                         int &steps) const {
        const long long tile = i >> 5;
        const int lane = (int)(i & 31);
        const int tiles_y = (P.height + FT_TILE_H - 1) / FT_TILE_H;
        const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
        const int xl = tx * FT_TILE_W + lane / FT_TILE_H;  // column in the rank's compact strip space
        const int y = ty * FT_TILE_H + lane % FT_TILE_H;
        const int strip = xl / P.strip_width;
        const int x = (strip * P.world + P.rank) * P.strip_width + (xl - strip * P.strip_width);
        if (xl >= P.local_cols || y >= P.height || x >= P.width) return false;
        slot = x * P.height + y;  // final position in FColor[X,Y]
        o = V3{cam.position[0], cam.position[1], cam.position[2]};
        d = pixel_direction(cam, x, y, P.max_size);
        len = P.length;
        eps = P.epsilon;
        steps = 0;
        if (len > 0.0f && d0 >= eps) {  // first march iteration with the shared Distance (see RenderJob)
            o = o + d * d0;
            len = len - d0;
            steps = 1;
        }
        return true;
    }
    __device__ __forceinline__ void store(slot_t slot, float r, float g, float b, int, int, unsigned) const {
        for (int k = 0; k < n_dst; k++) {
            float *out = dst[k] + 3ll * slot;
            out[0] = r;
            out[1] = g;
            out[2] = b;
        }
    }
    static constexpr bool kTileCost = true;
    __device__ __forceinline__ long long map_chunk(long long base) const { return hinted_chunk(P, base, n_items); }
    __device__ __forceinline__ long long draw_total() const { return hinted_total(P, n_items); }
    __device__ __forceinline__ void note_cost(long long base, unsigned cost) const {
        if (P.cost && base < n_items) P.cost[base >> 5] = (unsigned short)(cost < 65535u ? cost : 65535u);
    }
    // see RenderJob::store_tile; the tile's four columns are consecutive image columns when the strip width is a multiple of 4
    static constexpr bool kTileStore = true;
    __device__ __forceinline__ int *out_leaf_ptr() const { return nullptr; }
    __device__ __forceinline__ void store_tile(long long base_item, const float *sh, unsigned valid, int lane) const {
        const long long tile = base_item >> 5;
        const int tiles_y = (P.height + FT_TILE_H - 1) / FT_TILE_H;
        const int tx = (int)(tile / tiles_y), ty = (int)(tile - (long long)tx * tiles_y);
        const int xl0 = tx * FT_TILE_W, y0 = ty * FT_TILE_H;
        const int strip = xl0 / P.strip_width;
        const int x0 = (strip * P.world + P.rank) * P.strip_width + (xl0 - strip * P.strip_width);  // image column of the tile
        if (valid == 0xffffffffu && (P.height & 3) == 0 && (P.strip_width & 3) == 0) {
            const int col = lane / 6, part = lane - col * 6;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < 24) v = *reinterpret_cast<const float4 *>(sh + col * 24 + part * 4);
            const long long at = 3ll * ((long long)(x0 + col) * P.height + y0) + part * 4;
            for (int k = 0; k < n_dst; k++)
                if (lane < 24 && ((uintptr_t)dst[k] & 15) == 0) *reinterpret_cast<float4 *>(dst[k] + at) = v;
                else if (lane < 24) {
                    dst[k][at + 0] = v.x; dst[k][at + 1] = v.y; dst[k][at + 2] = v.z; dst[k][at + 3] = v.w;
                }
        } else if ((valid >> lane) & 1u) {
            const int xl = xl0 + lane / FT_TILE_H;
            const int st2 = xl / P.strip_width;
            const long long x = (long long)(st2 * P.world + P.rank) * P.strip_width + (xl - st2 * P.strip_width);
            const long long slot = x * P.height + (y0 + lane % FT_TILE_H);
            for (int k = 0; k < n_dst; k++) {
                float *out = dst[k] + 3 * slot;
                out[0] = sh[3 * lane + 0];
                out[1] = sh[3 * lane + 1];
                out[2] = sh[3 * lane + 2];
            }
        }
    }
};

// batched SdfScene.trace: 32-byte ft_ray records read as two LDG.128
struct RaysJob {
    const float4 *rays;
    long long n;
    float *out_rgb;
    int *out_leaf;
    int *out_steps;
    unsigned *out_shadow;

    typedef long long slot_t;
    __device__ __forceinline__ long long total() const { return n; }
    static constexpr bool kProbe = false;
    static constexpr bool kUniformEps = false;  // Epsilon is a field of every Ray record (Types.fs:16)
    __device__ __forceinline__ float uniform_eps() const { return 0.f; }
    __device__ __forceinline__ V3 probe_point() const { return V3{0.f, 0.f, 0.f}; }
    __device__ __forceinline__ bool load(long long i, float, V3 &o, V3 &d, float &len, float &eps, long long &slot,
                                         int &steps) const {
        steps = 0;
        const float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
        o = V3{a.x, a.y, a.z};
        d = V3{a.w, b.x, b.y};
        len = b.z;
        eps = b.w;
        slot = i;
        return true;
    }
    __device__ __forceinline__ void store(long long slot, float r, float g, float b, int leaf, int steps, unsigned shadow) const {
        if (out_rgb) {
            out_rgb[3 * slot + 0] = r;
            out_rgb[3 * slot + 1] = g;
            out_rgb[3 * slot + 2] = b;
        }
        if (out_leaf) out_leaf[slot] = leaf;
        if (out_steps) out_steps[slot] = steps;
        if (out_shadow) out_shadow[slot] = shadow;
    }
    static constexpr bool kTileStore = false;
    static constexpr bool kTileCost = false;
    __device__ __forceinline__ long long map_chunk(long long base) const { return base; }
    __device__ __forceinline__ long long draw_total() const { return n; }
    __device__ __forceinline__ void note_cost(long long, unsigned) const {}
    __device__ __forceinline__ int *out_leaf_ptr() const { return nullptr; }
    __device__ __forceinline__ void store_tile(long long, const float *, unsigned, int) const {}
};

// ESCAPE: resolve rays by the escape proof (identical colours / leaf / shadow mask; the step counters then
// count only the evaluations actually performed, so it is off when per-ray step counts are requested).
// CHAIN: the root form is a chain program (eval_chain) instead of a generic tree (eval_form).
template <bool COUNT, bool ESCAPE, bool CHAIN, class Job>
__global__ void __launch_bounds__(MB, FT_MINB) machine_kernel(const __grid_constant__ DevScene S, const __grid_constant__ Job job,
                                                     unsigned long long *__restrict__ work_counter,
                                                     unsigned long long *__restrict__ stats) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long total = job.total(), draw_total = job.draw_total();

    // lane state
    int phase = Job::kProbe ? PH_PROBE : PH_FETCH;
    float d0 = 0.f;
    typename Job::slot_t slot = 0;
    // `o` is ALWAYS the next evaluation point: the marching ray's origin, a normal probe point (the hit's leaf object is
    // resolved when the hit is detected, so the un-moved hit origin is not needed afterwards), or the job's probe point.
    // (A separate evaluation point `p` -- a copy of `o` or of a probe point, chosen at the head of every iteration -- cost
    // eleven instructions per iteration and three more live registers across the evaluator.)
    V3 o = Job::kProbe ? job.probe_point() : V3{0, 0, 0};
    V3 d{0, 0, 0}, q{0, 0, 0}, n{0, 0, 0}, lc{0, 0, 0};
    float len = 0.f, eps_lane = 1.f;  // per-ray Epsilon (RaysJob); render jobs use the uniform one
#define FT_EPS (Job::kUniformEps ? job.uniform_eps() : eps_lane)
    int steps = 0, primary_steps = 0, light = 0, leaf_obj = 0;
    // (measured and rejected: the leaf object, 27 bits, and the light counter, 5 bits, packed into one register between
    // iterations -- config 2 0.717 -> 0.727 ms, config 5 4K 2.473 -> 2.500; ptxas spills the same hot values either way)
    unsigned shadow_mask = 0;
    // the warp's current chunk (uniform)
    long long chunk_next = 0;
    int chunk_left = 0;
    bool exhausted = false;
    // per-lane totals for the stats flush
    // event counters live in shared memory (one row per warp, RED.shared on the rare transition paths)
    // FT_TILE_STORE: finished pixels wait in the warp's shared-memory row until the whole tile leaves at once (store_tile)
    constexpr bool TILE_STORE = FT_TILE_STORE && Job::kTileStore && FT_REFILL_MIN >= 32;
    // (Measured and rejected: dropping the lane's `slot` -- in whole-tile lockstep it can be formed again from the tile id when the
    // pixel is stored.  ptxas keeps ONE loop-carried value in the evaluator's result register and spills it around the body in
    // every iteration (STL at the head, LDL.LU at the end); without `slot` it picks another, and config 2 went 0.705 -> 0.716 ms.)
    __shared__ __align__(16) float sh_tile[TILE_STORE ? MB / 32 : 1][TILE_STORE ? 96 : 4];
    float *const wtile = sh_tile[TILE_STORE ? (threadIdx.x >> 5) : 0];
    long long tile_base = -1;  // the chunk in flight (uniform): staged pixels, cost record
    unsigned tile_t0 = 0u;     // SM clock when it was drawn: its cost is the time the warp spent on it, in units of 64 cycles
                               // (a loop-iteration counter was kept first: one more value incremented -- and spilled -- in every
                               // iteration of the march loop; the clock is read twice per tile)
    bool has_pix = false;
    __shared__ unsigned sh_stats[MB / 32][4];
    unsigned *const wstats = sh_stats[threadIdx.x >> 5];
    const unsigned sh_stats_base = (unsigned)__cvta_generic_to_shared(&sh_stats[0][0]);
    if (lane < 4) wstats[lane] = 0u;
    __syncwarp();
    Acc acc{0ull, 0ull};

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        // The warp draws 32-item chunks (one 4 x 8 pixel tile) from one global counter -- one atomic per chunk -- and deals the
        // chunk's items to idle lanes.  FT_REFILL_MIN idle lanes trigger a refill: 32 = whole tiles in lockstep (a new tile only
        // when every lane is idle), smaller values let finished lanes start on the next tile's pixels while their neighbours
        // are still marching (a warp then holds pixels of two or three neighbouring tiles).
        {
            unsigned idle = __ballot_sync(FULL, phase == PH_FETCH);
            if (idle == FULL || (__popc(idle) >= FT_REFILL_MIN && !(exhausted && chunk_left == 0))) {
                if ((TILE_STORE || Job::kTileCost) && idle == FULL && tile_base >= 0) {  // the previous tile is complete
                    if (TILE_STORE) {  // it leaves as a whole
                        __syncwarp();
                        job.store_tile(tile_base, wtile, __ballot_sync(FULL, has_pix), lane);
                        __syncwarp();
                        has_pix = false;
                    }
                    if (Job::kTileCost && lane == 0) job.note_cost(tile_base, ((unsigned)clock() - tile_t0) >> 6);
                    tile_base = -1;
                }
                while (idle) {
                    if (chunk_left == 0) {
                        if (exhausted) break;
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(work_counter, (unsigned long long)FT_CHUNK);
                        base = __shfl_sync(FULL, base, 0);
                        if ((long long)base >= draw_total) {
                            exhausted = true;
                            break;
                        }
                        // scheduling hint: the n-th chunk drawn is the n-th most expensive tile of the previous launch
                        chunk_next = job.map_chunk((long long)base);
                        chunk_left = FT_CHUNK;
                        if (TILE_STORE || Job::kTileCost) tile_base = chunk_next;
                        if (Job::kTileCost) tile_t0 = (unsigned)clock();
                    }
                    const int want = __popc(idle), take = want < chunk_left ? want : chunk_left;
                    const int r = __popc(idle & ((1u << lane) - 1u));
                    if (phase == PH_FETCH && r < take) {
                        const long long mine = chunk_next + r;
                        if (mine < total && job.load(mine, d0, o, d, len, eps_lane, slot, steps)) {
                            add_flops<COUNT>(acc, FLOPS_PRIMARY_RAY);
                            phase = PH_PRIMARY;
                        }
                    }
                    chunk_next += take;
                    chunk_left -= take;
                    idle = __ballot_sync(FULL, phase == PH_FETCH);  // padding items leave their lane idle: it draws again ...
                    if (FT_REFILL_MIN >= 32 && idle != FULL) break;  // ... but whole-tile lockstep never mixes two tiles in a warp
                }
                if (idle == FULL) {
                    if (exhausted) break;  // no work left anywhere
                    continue;
                }
            }
        }
        // ------------------------------------------------------------------ march pre-check (SdfForm.fs:94)
        // a ray whose Length is exhausted (or that hit the watchdog) ends as a miss without another evaluation
        int event = 0;  // 1 primary miss, 2 primary hit, 3 shadow finished (miss), 4 shadow finished (hit), 5 normal done
        uint32_t best_lo_hit = 0u;  // event 2: the arg-min candidate of the evaluation that detected the hit
        // the evaluation point is `o` in every phase (see the lane state)
        const V3 p = o;
        // chain scenes: one box test on the cell coordinate decides both "no escape pre-check" and "straight to the sub-grid"
        const V3 cc = CHAIN ? chain_coord(S, p) : V3{0.f, 0.f, 0.f};
        const bool fast = CHAIN && in_fast_region(S, cc);
        if (phase == PH_PRIMARY || phase == PH_SHADOW) {
            if (len <= 0.0f || steps >= S.step_cap) {
                if (!(len <= 0.0f)) warp_event(wstats, sh_stats_base, 2);
                event = phase == PH_PRIMARY ? 1 : 3;
            } else if (ESCAPE && S.esc_valid && !fast && outside_box(S.esc_lo, S.esc_hi, o) && ray_escapes(S, o, d, FT_EPS)) {
                warp_event(wstats, sh_stats_base, 3);
                event = phase == PH_PRIMARY ? 1 : 3;
            }
        }
        // ------------------------------------------------------------------ the one distance evaluation
        // Scheduling only (no arithmetic depends on it): positions outside the largest lookup's box scan whole
        // candidate lists (50-140 entries), positions inside it a handful.  The warp runs the two classes in
        // separate iterations so that one long scan does not hold 31 short ones hostage.
        const bool wants = phase > PH_FETCH && event == 0;
        const bool is_long = p.x < S.pg_lo[0] || p.y < S.pg_lo[1] || p.z < S.pg_lo[2] || p.x >= S.pg_hi[0] ||
                             p.y >= S.pg_hi[1] || p.z >= S.pg_hi[2];
#if FT_LONG_VOTE
        const unsigned m_long = __ballot_sync(FULL, wants && is_long), m_short = __ballot_sync(FULL, wants && !is_long);
        const bool run_long = m_short == 0u || __popc(m_long) >= FT_LONG_BATCH;
        if (wants && is_long == run_long) {
#else
        if (wants) {
#endif
            const unsigned long long f0 = acc.flops;
            uint32_t best_lo = 0u;  // the arg-min candidate of this evaluation (chain scenes)
            const float dist = CHAIN ? eval_chain_c<COUNT>(S, p, cc, fast, acc, best_lo) : eval_form<COUNT>(S, S.root_word, p, acc);
            if (COUNT) {
                acc.evals++;
                if (is_long) {
                    acc.long_flops += acc.flops - f0;
                    acc.long_evals++;
                }
            }
            best_lo_hit = best_lo;
            if (phase <= PH_SHADOW) {  // SdfForm.fs:97-104
                steps++;
                if (COUNT) acc.steps++;
                if (dist < FT_EPS) {
                    event = phase == PH_PRIMARY ? 2 : 4;  // (a primary hit resolves its leaf object from `best_lo` below)
                } else {
                    o = o + d * dist;  // Ray.move
                    len = len - dist;
                    add_flops<COUNT>(acc, FLOPS_MOVE);
                }
            } else if (Job::kProbe && phase == PH_PROBE) {
                d0 = dist;
                phase = PH_FETCH;
            } else if (phase == PH_N0) {  // the probe points of SdfForm.fs:107-112: q + h e_x, q + h e_y, q + h e_z, q
                n.x = dist;
                o.x = q.x;
                o.y = q.y + FT_EPS * 0.125f;
                phase = PH_N1;
            } else if (phase == PH_N1) {
                n.y = dist;
                o.y = q.y;
                o.z = q.z + FT_EPS * 0.125f;
                phase = PH_N2;
            } else if (phase == PH_N2) {
                n.z = dist;
                o.z = q.z;
                phase = PH_N3;
            } else {
                n = normalize(n - V3{dist, dist, dist});  // SdfForm.fs:107-112
                add_flops<COUNT>(acc, FLOPS_NORMAL);
                event = 5;
            }
        }
        // ------------------------------------------------------------------ transitions (once or a few times per ray)
        if (event != 0) {
            bool next_light = false, finish = false;
            float r = S.bg[0], g = S.bg[1], b = S.bg[2];
            int leaf = -1;
            if (event == 1) {  // SdfScene.fs:10
                primary_steps = steps;
                shadow_mask = 0;
                finish = true;
            } else if (event == 2) {  // primary hit: SdfObject.fs:70-73
                primary_steps = steps;
                warp_event(wstats, sh_stats_base, 0);
                // Material.Color at the un-moved hit origin `o` (SdfObject.fs:75-77): the leaf object is resolved here, before
                // the normal (the reference takes the normal first, SdfObject.fs:70-77; both are pure functions of the hit), from
                // the arg-min candidate of the evaluation that detected the hit or by the arg-min scan
                leaf_obj = (FT_CHAIN_MAT && CHAIN && S.chain_mat_valid) ? chain_leaf_from_best(S, o, best_lo_hit)
                                                                         : material_leaf<COUNT>(S, o, acc);
                q = o + d * (-FT_EPS);  // Ray.get (-epsilon): the normal's probe point and result.Ray's origin
                o = q;
                o.x = q.x + FT_EPS * 0.125f;  // h = epsilon / 8, SdfForm.fs:115
                phase = PH_N0;
            } else if (event == 5) {
                lc = V3{S.bg[0], S.bg[1], S.bg[2]};  // SdfScene.fs:12
                shadow_mask = 0;
                light = 0;
                next_light = true;
            } else {  // shadow march finished
                if (event == 3) {  // SdfScene.fs:21-23 no shadow: lightColor + intensity * cos
                    // intensity * cos is recomputed from the light record (same inputs, same operations) instead of being
                    // carried in three registers through the whole shadow march
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    V3 intensity = xyz(lrgb);
                    float lcos;
                    if (__float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL) {
                        lcos = dot(n, xyz(lv));
                    } else {
                        lcos = dot(n, normalize(xyz(lv) - q));
                        intensity = xyz(lrgb) / length2(xyz(lv) - q);
                    }
                    lc = lc + intensity * lcos;
                    shadow_mask |= 1u << light;
                }
                light++;
                next_light = true;
            }
            if (next_light) {
                finish = true;
                for (; light < S.n_lights; light++) {  // SdfScene.fs:13-19
                    const float4 lv = ld4(S.lights + 2 * light), lrgb = ld4(S.lights + 2 * light + 1);
                    const bool directional = __float_as_int(lv.w) == FT_LIGHT_DIRECTIONAL;
                    const V3 ldir = directional ? xyz(lv) : normalize(xyz(lv) - q);  // SdfLight.fs:9 / :25
                    const float lcos = dot(n, ldir);
                    add_flops<COUNT>(acc, FLOPS_LIGHT);
                    if (lcos > 0.0f) {
                        o = q;
                        if (directional) {  // SdfLight.fs:10-20
                            d = xyz(lv);
                            len = 1000.0f;
                        } else {  // SdfLight.fs:26-41
                            const V3 diff = xyz(lv) - q;
                            const float d2 = length2(diff);
                            len = sqrtf(d2);
                            d = diff / d2;
                            add_flops<COUNT>(acc, FLOPS_POINT_LIGHT);
                        }
                        steps = 0;
                        warp_event(wstats, sh_stats_base, 1);
                        phase = PH_SHADOW;
                        finish = false;
                        break;
                    }
                }
                if (finish) {  // SdfScene.fs:28
                    const int4 lrec = __ldg(S.objects + leaf_obj);
                    const float4 col = ld4(S.materials + lrec.z);
                    const float pi_inv = 1.0f / 3.14159274f;
                    const V3 c = xyz(col) * (lc * pi_inv);
                    add_flops<COUNT>(acc, FLOPS_FINAL);
                    r = c.x; g = c.y; b = c.z;
                    leaf = lrec.w;
                }
            }
            if (finish) {
                if (TILE_STORE) {
                    wtile[3 * lane + 0] = r;
                    wtile[3 * lane + 1] = g;
                    wtile[3 * lane + 2] = b;
                    has_pix = true;
                    if (job.out_leaf_ptr()) job.out_leaf_ptr()[slot] = leaf;
                } else {
                    job.store(slot, r, g, b, leaf, primary_steps, shadow_mask);
                }
                phase = PH_FETCH;
            }
        }
    }

    // ---------------------------------------------------------------------- stats
    __syncwarp();
    const unsigned hits = wstats[0], shadows = wstats[1], capped = wstats[2], escaped = wstats[3];
    unsigned long long st = acc.steps, fl = acc.flops, lf = acc.long_flops, le = acc.long_evals, ev = acc.evals;
    if (COUNT) {
        for (int k = 16; k > 0; k >>= 1) {
            st += __shfl_xor_sync(FULL, st, k);
            fl += __shfl_xor_sync(FULL, fl, k);
            lf += __shfl_xor_sync(FULL, lf, k);
            le += __shfl_xor_sync(FULL, le, k);
            ev += __shfl_xor_sync(FULL, ev, k);
        }
    }
    if (lane == 0 && stats) {  // stats == nullptr: an asynchronous launch whose caller asked for no counters
        const unsigned w = (blockIdx.x * (MB / 32) + (threadIdx.x >> 5)) & (FT_STAT_SLOTS - 1);
        unsigned long long *s = stats + w * FT_STAT_STRIDE;
        if (hits) atomicAdd(s + FT_STAT_HITS, (unsigned long long)hits);
        if (shadows) atomicAdd(s + FT_STAT_SHADOW, (unsigned long long)shadows);
        if (capped) atomicAdd(s + FT_STAT_CAPPED, (unsigned long long)capped);
        if (escaped) atomicAdd(s + FT_STAT_ESCAPED, (unsigned long long)escaped);
        if (COUNT) {
            atomicAdd(s + FT_STAT_STEPS, st);
            atomicAdd(s + FT_STAT_FLOPS, fl);
            atomicAdd(s + FT_STAT_LONG_FLOPS, lf);
            atomicAdd(s + FT_STAT_LONG_EVALS, le);
            atomicAdd(s + FT_STAT_EVALS, ev);
        }
    }
}

#undef FT_EPS
}  // namespace FT_NS
