// ft_api.cu -- the C ABI of include/fraytracer_b200.h: scene handles, device memory, streams, stats.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "scene_compiler.h"
#include "spectral.h"
#include "cloud.h"
#include "subgrid.h"

namespace {

constexpr unsigned kCounterRing = 256;

// Three builds of the hot-path kernels: strict without the extension cone primitive (the reference's own primitives: the
// default and benchmarked path -- even an untaken cone branch inside the one evaluator call site cost scenes without cones
// 4 %, measured), strict with it (chosen only when the scene holds a cone), and the contracted (FMA) build.
#define FT_DISPATCH(FAST, FN, ...)                                                           \
    ((FAST) ? ft_fast::FN(__VA_ARGS__)                                                       \
            : (scene->host.has_cone ? ft_strict_cone::FN(__VA_ARGS__) : ft_strict::FN(__VA_ARGS__)))

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return FT_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FT_ERR_OUT_OF_MEMORY;
    return FT_ERR_CUDA;
}
#define CU(call)                                           \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
cudaError_t upload(const std::vector<T> &v, const T **out, int64_t *bytes) {
    T *d = nullptr;
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&d, n);
    if (e != cudaSuccess) return e;
    if (!v.empty()) {
        e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return e;
        }
    }
    *out = d;
    *bytes += (int64_t)n;
    return cudaSuccess;
}

}  // namespace

// Scheduling hint of one launch shape (frame size, strip partition, band, kind of destination): the per-tile costs its launches
// record and the tile orders derived from them (tile_rank.cu).  Purely a scheduling matter: pixels do not depend on the order
// tiles are drawn in (tests render with and without hints).  Frame N of a shape draws its tiles in the order ranked from
// frame N - 2's costs: the ranking of frame N - 1 runs on the scene's side stream while frame N is being rendered, so that
// back-to-back launches never wait for a sort.
struct TileHint {
    int key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned n_tiles = 0;
    unsigned short *d_cost = nullptr;
    TileRank *rank = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};  // order[parity] has been written
    cudaEvent_t snapped = nullptr;              // the last frame's costs have been copied out of d_cost
    bool valid[2] = {false, false};
    unsigned long long frames = 0, stamp = 0;
    int mode = 0;  // FT_RANK_*
    void release() {
        for (cudaEvent_t e : {ready[0], ready[1], snapped})
            if (e) {
                cudaEventSynchronize(e);
                cudaEventDestroy(e);
            }
        if (d_cost) cudaFree(d_cost);
        tile_rank_destroy(rank);
        *this = TileHint();
    }
};
constexpr int kTileHints = 16;

struct ft_scene {
    int device = 0;
    int sm_count = 0;
    ftc::CompiledScene host;
    DevScene dev{};
    int64_t device_bytes = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long *d_stats = nullptr, *h_stats = nullptr;
    // sub_debug bit 0 ("distrust every certificate", tests): the sub-grid is detached from the kernel parameters, so every
    // evaluation runs the reference's sequential scan; the built state is kept to re-attach it
    int32_t sub_valid_built = 0;
    float fast_lo_built[3] = {0.f, 0.f, 0.f}, fast_hi_built[3] = {0.f, 0.f, 0.f};
    void apply_sub_debug() {
        const bool detach = (dev.sub_debug & 1) != 0;
        dev.sub_valid = detach ? 0 : sub_valid_built;
        for (int k = 0; k < 3; k++) {
            dev.fast_lo[k] = detach ? std::numeric_limits<float>::infinity() : fast_lo_built[k];
            dev.fast_hi[k] = detach ? -std::numeric_limits<float>::infinity() : fast_hi_built[k];
        }
        // fine-cell units (DevScene::chain_g2s): exact scalings by a power of two
        const float scale = dev.sub_valid ? dev.sub_scale : 1.0f;
        dev.chain_g2s = float4{dev.chain_g2.x * scale, dev.chain_g2.y * scale, dev.chain_g2.z * scale, 0.f};
        for (int k = 0; k < 3; k++) {
            dev.fast_lo_s[k] = dev.fast_lo[k] * scale;
            dev.fast_hi_s[k] = dev.fast_hi[k] * scale;
        }
    }
    // ring of work counters for the persistent kernels: every launch zeroes and uses the next one, so that
    // launches in flight on different streams never share a counter
    unsigned long long *d_counters = nullptr;
    unsigned counter_next = 0;
    // scratch for the host-pointer entry points, grown on demand
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // second scratch (spectral extension: path queues and slots, sized after wave 0) and its wave counters
    int *d_cloud_perm = nullptr;  // volumetric extension: the noise tables of the last ft_render_cloud call
    double *d_cloud_grad = nullptr;
    void *d_scratch2 = nullptr;
    size_t scratch2_bytes = 0;
    unsigned *d_spec_counts = nullptr, *h_spec_counts = nullptr;
    SubGrid sub;  // nearest-candidate sub-grid of a chain scene's spatial lookup (subgrid_build.cu)
    TileHint hints[kTileHints];
    cudaStream_t hint_stream = nullptr;  // the tile rankings are computed here, beside the caller's stream
    unsigned long long hint_clock = 0;
    bool hints_enabled = true;
    bool ext_count = false;  // ft_scene_set_debug bit 3
    // destinations behind PCIe (mapped host memory): FT_RANK_TIMED by default (tile_rank.cu); 0 = plain LPT, 1 = LPT read from
    // both ends alternately (RenderParams::order_mode) -- experiments
    // (re-measured after the round's kernel work, ft_render config 2 / config 5 at 4K, ms: split 110 0.905 / 3.32, 150 0.876 / 2.96..3.14,
    // 200 0.881 / 2.90, 300 0.899 / 2.77; shift 1 0.866..0.889 / 2.88..3.2, shift 2 0.931 / 2.82: nothing beats the defaults on both)
    int hint_pcie_mode = FT_RANK_TIMED, hint_pcie_shift = 0, hint_split_pct = 150;
    std::mutex mu;

    cudaError_t reserve(size_t bytes) {
        if (bytes <= scratch_bytes) return cudaSuccess;
        if (d_scratch) cudaFree(d_scratch);
        d_scratch = nullptr;
        scratch_bytes = 0;
        cudaError_t e = cudaMalloc(&d_scratch, bytes);
        if (e == cudaSuccess) scratch_bytes = bytes;
        return e;
    }
    cudaError_t reserve2(size_t bytes) {
        if (bytes <= scratch2_bytes) return cudaSuccess;
        if (d_scratch2) cudaFree(d_scratch2);
        d_scratch2 = nullptr;
        scratch2_bytes = 0;
        cudaError_t e = cudaMalloc(&d_scratch2, bytes);
        if (e == cudaSuccess) scratch2_bytes = bytes;
        return e;
    }
    void release() {
        const void *ptrs[] = {dev.nodes, dev.cells, dev.entries, dev.links,
                              dev.objects, dev.obj_links, dev.materials, dev.lights};
        for (const void *p : ptrs)
            if (p) cudaFree(const_cast<void *>(p));
        if (hint_stream) cudaStreamSynchronize(hint_stream);
        for (auto &h : hints) h.release();
        if (hint_stream) cudaStreamDestroy(hint_stream);
        if (sub.hdr) cudaFree(sub.hdr);
        if (sub.idx) cudaFree(sub.idx);
        if (d_stats) cudaFree(d_stats);
        if (d_counters) cudaFree(d_counters);
        if (h_stats) cudaFreeHost(h_stats);
        if (d_scratch) cudaFree(d_scratch);
        if (d_scratch2) cudaFree(d_scratch2);
        if (d_spec_counts) cudaFree(d_spec_counts);
        if (d_cloud_perm) cudaFree(d_cloud_perm);
        if (d_cloud_grad) cudaFree(d_cloud_grad);
        if (h_spec_counts) cudaFreeHost(h_spec_counts);
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

namespace {

constexpr size_t kStatWords = (size_t)FT_STAT_SLOTS * FT_STAT_STRIDE;

cudaError_t fresh_counter(ft_scene *scene, cudaStream_t stream, unsigned long long **out) {
    unsigned long long *c = scene->d_counters + (size_t)(scene->counter_next++ % kCounterRing) * 16;
    *out = c;
    return cudaMemsetAsync(c, 0, sizeof(unsigned long long), stream);
}

void sum_stats(const unsigned long long *h, ft_render_stats *st) {
    unsigned long long hits = 0, shadow = 0, capped = 0, steps = 0, flops = 0, escaped = 0;
    for (int s = 0; s < FT_STAT_SLOTS; s++) {
        const unsigned long long *p = h + (size_t)s * FT_STAT_STRIDE;
        hits += p[FT_STAT_HITS];
        shadow += p[FT_STAT_SHADOW];
        capped += p[FT_STAT_CAPPED];
        steps += p[FT_STAT_STEPS];
        flops += p[FT_STAT_FLOPS];
        escaped += p[FT_STAT_ESCAPED];
    }
    if (getenv("FT_PRINT_LONG")) {  // experiments: share of the work done outside the largest grid's box
        unsigned long long lf = 0, le = 0, ev = 0;
        for (int k = 0; k < FT_STAT_SLOTS; k++) {
            lf += h[(size_t)k * FT_STAT_STRIDE + FT_STAT_LONG_FLOPS];
            le += h[(size_t)k * FT_STAT_STRIDE + FT_STAT_LONG_EVALS];
            ev += h[(size_t)k * FT_STAT_STRIDE + FT_STAT_EVALS];
        }
        fprintf(stderr, "[ft] evaluations %llu (outside the grid box %llu), flops %llu (outside %llu)\n", ev, le, flops, lf);
    }
    st->escaped_rays = (int64_t)escaped;
    st->primary_hits = (int64_t)hits;
    st->shadow_rays = (int64_t)shadow;
    st->capped_rays = (int64_t)capped;
    st->march_steps = (int64_t)steps;
    st->eval_flops = (int64_t)flops;
}

int check_frame(int64_t width, int64_t height) {
    if (width <= 0 || height <= 0 || width * height >= ((int64_t)1 << 29))
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid image size (a frame holds at most 2^29 pixels)");
    return FT_OK;
}

int check_ray_params(float epsilon, float length) {
    if (!(epsilon > 0.0f))
        return fail(FT_ERR_INVALID_ARGUMENT,
                    "epsilon must be > 0: with epsilon <= 0 the reference's march loop (SdfForm.fs:93-104) need not "
                    "terminate");
    if (std::isnan(length)) return fail(FT_ERR_INVALID_ARGUMENT, "length is NaN");
    return FT_OK;
}

RenderParams make_params(int width, int height, float eps, float length, int strip_width, int rank, int world) {
    RenderParams P;
    P.width = width;
    P.height = height;
    P.strip_width = strip_width;
    P.rank = rank;
    P.world = world;
    P.x_offset = 0;
    P.xl_begin = 0;
    P.order = nullptr;
    P.cost = nullptr;
    P.order_mode = 0;
    P.order_shift = 0;
    const int strips = (width + strip_width - 1) / strip_width;
    const int spr = (strips + world - 1) / world;
    P.local_cols = world == 1 ? width : spr * strip_width;
    P.epsilon = eps;
    P.length = length;
    P.max_size = (float)std::max(width, height);  // Image.fs:18
    return P;
}


// The hint slot of a launch shape (least recently used slot is recycled); null when hints are off or allocation fails.
// Attaches the order (if one exists yet) and the cost buffer to P and orders `s` after the slot's earlier users.
TileHint *hint_attach(ft_scene *scene, RenderParams &P, int tag, cudaStream_t s, bool behind_pcie = false) {
    if (!scene->hints_enabled) return nullptr;
    const long long tiles = (long long)((P.local_cols + 3) / 4) * ((P.height + 7) / 8);
    if (tiles <= 0 || tiles > (1ll << 26)) return nullptr;
    const int key[8] = {P.width, P.height, P.local_cols, P.world > 1 ? P.strip_width : 0, P.rank, P.world, P.x_offset * 2 + tag, P.xl_begin};
    TileHint *slot = nullptr, *lru = &scene->hints[0];
    for (auto &h : scene->hints) {
        if (h.n_tiles && std::memcmp(h.key, key, sizeof(key)) == 0) slot = &h;
        if (h.stamp < lru->stamp) lru = &h;
    }
    if (!slot) {
        slot = lru;
        slot->release();  // waits for its last users
        std::memcpy(slot->key, key, sizeof(key));
        slot->n_tiles = (unsigned)tiles;
        slot->mode = behind_pcie ? scene->hint_pcie_mode : FT_RANK_LPT;
        if (cudaMalloc(&slot->d_cost, sizeof(unsigned short) * (size_t)tiles) != cudaSuccess ||
            tile_rank_create((unsigned)tiles, behind_pcie ? scene->hint_pcie_shift : 0, s, &slot->rank) != cudaSuccess ||
            cudaEventCreateWithFlags(&slot->ready[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&slot->ready[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&slot->snapped, cudaEventDisableTiming) != cudaSuccess ||
            cudaMemsetAsync(slot->d_cost, 0, sizeof(unsigned short) * (size_t)tiles, s) != cudaSuccess) {
            cudaGetLastError();
            slot->release();
            return nullptr;
        }
    }
    const int parity = (int)(slot->frames & 1ull);
    // the kernel overwrites d_cost (the previous frame's costs must have been copied out) and reads order[parity]
    if ((slot->frames > 0 && cudaStreamWaitEvent(s, slot->snapped, 0) != cudaSuccess) ||
        (slot->valid[parity] && cudaStreamWaitEvent(s, slot->ready[parity], 0) != cudaSuccess)) {
        cudaGetLastError();
        return nullptr;
    }
    slot->stamp = ++scene->hint_clock;
    P.order = slot->valid[parity] ? tile_rank_order(slot->rank, parity) : nullptr;
    P.cost = slot->d_cost;
    P.order_mode = slot->mode == 1 ? 1 : 0;
    P.order_shift = tile_rank_shift(slot->rank);
    return slot;
}

// after the render kernel has been enqueued on `s`: keep its costs and rank them on the scene's side stream
void hint_finish(ft_scene *scene, TileHint *h, cudaStream_t s) {
    if (!h) return;
    const int parity = (int)(h->frames & 1ull);
    h->frames++;
    cudaStream_t side = scene->hint_stream ? scene->hint_stream : s;
    if (tile_rank_snapshot(h->rank, h->d_cost, parity, s) == cudaSuccess && cudaEventRecord(h->snapped, s) == cudaSuccess &&
        (side == s || cudaStreamWaitEvent(side, h->snapped, 0) == cudaSuccess) &&
        tile_rank_run(h->rank, parity, h->mode == FT_RANK_TIMED ? FT_RANK_TIMED : FT_RANK_LPT, scene->hint_split_pct, side) == cudaSuccess &&
        cudaEventRecord(h->ready[parity], side) == cudaSuccess) {
        h->valid[parity] = true;
    } else {
        cudaGetLastError();
        h->valid[parity] = false;
    }
}

// pixels of the frame the strips of `T` cover (padding columns of the last strip excluded)
int64_t owned_pixels(const StripTarget &T, int width, int height) {
    if (T.strip_width == 0) {  // tiles dealt round-robin
        const int64_t tiles_y = (height + 7) / 8, total = (int64_t)((width + 3) / 4) * tiles_y;
        int64_t n = 0;
        for (int64_t t = T.rank; t < total; t += T.world) {
            const int x0 = (int)(t / tiles_y) * 4, y0 = (int)(t % tiles_y) * 8;
            n += (int64_t)std::min(4, width - x0) * std::min(8, height - y0);
        }
        return n;
    }
    int64_t cols = 0;
    for (int xl = 0; xl < T.local_cols; xl++) {
        const int strip = xl / T.strip_width;
        cols += (strip * T.world + T.rank) * T.strip_width + (xl - strip * T.strip_width) < width;
    }
    return cols * height;
}

// 4 x 8 pixel tiles this rank renders
long long local_tiles(const StripTarget &T, int width, int height) {
    const long long tiles_y = (height + 7) / 8;
    if (T.strip_width == 0) {
        const long long total = (long long)((width + 3) / 4) * tiles_y;
        return total > T.rank ? (total - T.rank + T.world - 1) / T.world : 0;
    }
    return (long long)((T.local_cols + 3) / 4) * tiles_y;
}

StripTarget whole_frame(int width) {
    StripTarget T{};
    T.strip_width = 8;
    T.rank = 0;
    T.world = 1;
    T.local_cols = width;
    return T;
}

int make_target(int width, int strip_width, int rank, int world, float *const *d_dst, int n_dst, StripTarget *out) {
    if (strip_width < 0 || world <= 0 || rank < 0 || rank >= world || !d_dst || n_dst < 1 || n_dst > FT_MAX_DST)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid strip partition or destination count");
    StripTarget T{};
    T.strip_width = strip_width;
    T.rank = rank;
    T.world = world;
    if (strip_width > 0) {
        const int strips = (width + strip_width - 1) / strip_width, spr = (strips + world - 1) / world;
        T.local_cols = world == 1 ? width : spr * strip_width;
    } else {
        T.local_cols = width;  // tiles dealt round-robin: sizes below come from local_tiles()
    }
    T.n_dst = n_dst;
    for (int k = 0; k < n_dst; k++) {
        if (!d_dst[k]) return fail(FT_ERR_INVALID_ARGUMENT, "null destination framebuffer");
        T.dst[k] = d_dst[k];
    }
    *out = T;
    return FT_OK;
}

// Spectral wavefront pipeline for the strips `T` describes.  host_out != NULL: the frame is rendered into the scene's scratch
// image (T.dst[0] is set here) and copied to the host; otherwise every pixel goes to T's destination framebuffers.
int spectral_core(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon, float length,
                  int32_t n_wavelengths, int32_t max_bounces, uint64_t seed, StripTarget T, cudaStream_t s, float *out_rgb,
                  ft_render_stats *stats) {
    if (!scene || !camera) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (width <= 0 || height <= 0 || (int64_t)width * height >= ((int64_t)1 << 31))
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid image size");
    if (n_wavelengths < 1 || n_wavelengths > 64) return fail(FT_ERR_INVALID_ARGUMENT, "n_wavelengths must be in 1..64");
    if (max_bounces < 0 || max_bounces > 256) return fail(FT_ERR_INVALID_ARGUMENT, "max_bounces must be in 0..256");
    int st = check_ray_params(epsilon, length);
    if (st != FT_OK) return st;
    if (scene->host.n_lights > FT_SPECTRAL_MAX_LIGHTS)
        return fail(FT_ERR_INVALID_ARGUMENT, "ft_render_spectral supports at most FT_SPECTRAL_MAX_LIGHTS lights");
    if (scene->host.has_cloud)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_CLOUD (volumetric extension): render it with ft_render_cloud");
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t npix = (size_t)width * height, npix_local = (size_t)local_tiles(T, width, height) * 32;
    const size_t rgb_bytes = out_rgb ? (npix * 3 * sizeof(float) + 255) & ~(size_t)255 : 0;
    CU(scene->reserve(rgb_bytes + npix_local * 3 * sizeof(float4)));  // (image +) worst-case hit queue of this rank's strips
    if (!scene->d_spec_counts) {
        CU(cudaMalloc(&scene->d_spec_counts, 1024 * sizeof(unsigned)));
        CU(cudaMallocHost(&scene->h_spec_counts, 1024 * sizeof(unsigned)));
    }
    SpectralParams P{};
    P.cam = *camera;
    P.width = width;
    P.height = height;
    P.epsilon = epsilon;
    P.length = length;
    P.max_size = (float)std::max(width, height);
    P.nl = n_wavelengths;
    P.max_bounces = max_bounces;
    P.seed_lo = (uint32_t)seed;
    P.seed_hi = (uint32_t)(seed >> 32);
    P.n_items = local_tiles(T, width, height) * 32;
    if (out_rgb) {
        T.n_dst = 1;
        T.dst[0] = (float *)scene->d_scratch;
    }
    P.T = T;
    P.hit_q = (float4 *)((char *)scene->d_scratch + rgb_bytes);
    unsigned *counts = scene->d_spec_counts;  // [0] glass hits, [2 w], [2 w + 1] live paths of either class entering wave w
    P.hit_count = counts;
    P.stats = scene->d_stats;
    CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(counts, 0, 1024 * sizeof(unsigned), s));
    CU(cudaEventRecord(scene->ev[0], s));
    int launches = 0;
    // ---- wave 0: primary rays
    P.wave = 0;
    CU(fresh_counter(scene, s, &P.work_counter));
    const bool count = scene->ext_count;
    cudaError_t e = spectral_launch_primary(scene->dev, P, count, scene->sm_count, s);
    if (e != cudaSuccess) return cuda_fail(e, "spectral primary kernel launch");
    launches++;
    CU(cudaMemcpyAsync(scene->h_spec_counts, counts, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));  // the one host round trip: the path queues are sized by the glass hit count
    const unsigned n_hits = scene->h_spec_counts[0];
    const unsigned long long n_paths0 = (unsigned long long)n_hits * (unsigned)n_wavelengths;
    if (n_paths0 >= (1ull << 31)) return fail(FT_ERR_INVALID_ARGUMENT, "too many spectral paths for one call (>= 2^31)");
    if (n_hits > 0) {
        const size_t slot_bytes = (n_paths0 * sizeof(float) + 255) & ~(size_t)255;
        const size_t q_bytes = max_bounces > 0 ? n_paths0 * 3 * sizeof(float4) : 0;
        CU(scene->reserve2(slot_bytes + 2 * q_bytes));
        P.slots = (float *)scene->d_scratch2;
        float4 *qa = (float4 *)((char *)scene->d_scratch2 + slot_bytes), *qb = (float4 *)((char *)qa + q_bytes);
        // ---- the material stage of wave 0
        P.out_q = qa;
        P.out_count = counts + 2;
        P.q_cap = (unsigned)n_paths0;
        P.split = getenv("FT_SPEC_NO_SPLIT") == nullptr;  // experiments only
        e = spectral_launch_spawn(scene->dev, P, n_hits, s);
        if (e != cudaSuccess) return cuda_fail(e, "spectral spawn kernel launch");
        launches++;
        if (max_bounces > 0) {
            // ---- waves 1 .. max_bounces (a wave whose queue is empty exits at once)
            for (int w = 1; w <= max_bounces; w++) {
                P.wave = w;
                P.in_q = (w & 1) ? qa : qb;
                P.out_q = (w & 1) ? qb : qa;
                P.in_count = counts + 2 * w;
                P.out_count = counts + 2 * w + 2;
                CU(fresh_counter(scene, s, &P.work_counter));
                e = spectral_launch_wave(scene->dev, P, (long long)n_paths0, count, scene->sm_count, s);
                if (e != cudaSuccess) return cuda_fail(e, "spectral wave kernel launch");
                launches++;
            }
        }
        e = spectral_launch_resolve(P, n_hits, s);
        if (e != cudaSuccess) return cuda_fail(e, "spectral resolve kernel launch");
        launches++;
    }
    CU(cudaEventRecord(scene->ev[1], s));
    if (out_rgb) CU(cudaMemcpyAsync(out_rgb, T.dst[0], npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(scene->ev[2], s));
    CU(cudaStreamSynchronize(s));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        sum_stats(scene->h_stats, stats);
        unsigned long long paths = 0;
        for (int k = 0; k < FT_STAT_SLOTS; k++) paths += scene->h_stats[(size_t)k * FT_STAT_STRIDE + FT_STAT_PATHS];
        stats->primary_rays = owned_pixels(T, width, height);
        stats->shadow_rays += (int64_t)paths;  // every secondary ray: shadow rays + continuation segments
        stats->reserved = (int32_t)n_hits;     // pixels whose primary ray hit glass
        if (count)  // the material stage (60 per path, FLOPS_GLASS) and the per-pixel wavelength sums (9 per wavelength + 3)
            stats->eval_flops += (int64_t)n_paths0 * (max_bounces > 0 ? 60 : 0) + (int64_t)n_hits * (9 * n_wavelengths + 3);
        stats->launches = launches;
        cudaEventElapsedTime(&stats->kernel_ms, scene->ev[0], scene->ev[1]);
        cudaEventElapsedTime(&stats->total_ms, scene->ev[0], scene->ev[2]);
    }
    return FT_OK;
}


// Volumetric cloud kernel for the strips `T` describes; host_out as in spectral_core.
int cloud_core(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon, float length,
               const ft_cloud_params *cloud, int32_t samples_per_pixel, uint64_t seed, StripTarget T, cudaStream_t s,
               float *out_rgb, ft_render_stats *stats) {
    if (!scene || !camera || !cloud || !cloud->perm || !cloud->gradients) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (width <= 0 || height <= 0 || (int64_t)width * height >= ((int64_t)1 << 31))
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid image size");
    if (samples_per_pixel < 1 || samples_per_pixel > 65536) return fail(FT_ERR_INVALID_ARGUMENT, "samples_per_pixel must be in 1..65536");
    if (!(cloud->step > 0.0f) || !(cloud->fade > 0.0f)) return fail(FT_ERR_INVALID_ARGUMENT, "cloud step and fade must be > 0");
    for (int i = 0; i < 256; i++)
        if (cloud->perm[i] < 0 || cloud->perm[i] > 255) return fail(FT_ERR_INVALID_ARGUMENT, "cloud perm entries must be in 0..255");
    int st = check_ray_params(epsilon, length);
    if (st != FT_OK) return st;
    if (scene->host.has_glass)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_GLASS (spectral extension): render it with ft_render_spectral");
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t npix = (size_t)width * height;
    if (out_rgb) {
        CU(scene->reserve(npix * 3 * sizeof(float)));
        T.n_dst = 1;
        T.dst[0] = (float *)scene->d_scratch;
    }
    if (!scene->d_cloud_perm) {
        CU(cudaMalloc(&scene->d_cloud_perm, 256 * sizeof(int)));
        CU(cudaMalloc(&scene->d_cloud_grad, 768 * sizeof(double)));
    }
    CU(cudaMemcpyAsync(scene->d_cloud_perm, cloud->perm, 256 * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(scene->d_cloud_grad, cloud->gradients, 768 * sizeof(double), cudaMemcpyHostToDevice, s));
    CloudParams P{};
    P.cam = *camera;
    P.width = width;
    P.height = height;
    P.epsilon = epsilon;
    P.length = length;
    P.max_size = (float)std::max(width, height);
    P.spp = samples_per_pixel;
    P.seed_lo = (uint32_t)seed;
    P.seed_hi = (uint32_t)(seed >> 32);
    P.n_tiles = local_tiles(T, width, height);
    CU(scene->reserve2((size_t)std::max<long long>(P.n_tiles, 1) * 32 * FT_CLOUD_SAMPLE_GROUPS * 3 * sizeof(float)));
    P.partials = (float *)scene->d_scratch2;
    P.density = cloud->density;
    P.frequency = cloud->frequency;
    P.step = cloud->step;
    P.fade = cloud->fade;
    P.bias = cloud->bias;
    P.gain = cloud->gain;
    P.perm = scene->d_cloud_perm;
    P.gradients = scene->d_cloud_grad;
    P.T = T;
    P.stats = scene->d_stats;
    CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    CU(fresh_counter(scene, s, &P.work_counter));
    CU(cudaEventRecord(scene->ev[0], s));
    cudaError_t e = cloud_launch(scene->dev, P, scene->ext_count, scene->sm_count, s);
    if (e != cudaSuccess) return cuda_fail(e, "cloud kernel launch");
    CU(cudaEventRecord(scene->ev[1], s));
    if (out_rgb) CU(cudaMemcpyAsync(out_rgb, T.dst[0], npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(scene->ev[2], s));
    CU(cudaStreamSynchronize(s));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        sum_stats(scene->h_stats, stats);
        unsigned long long scattered = 0;
        for (int k = 0; k < FT_STAT_SLOTS; k++) scattered += scene->h_stats[(size_t)k * FT_STAT_STRIDE + FT_STAT_PATHS];
        stats->primary_rays = owned_pixels(T, width, height) * samples_per_pixel;
        stats->reserved = (int32_t)std::min<unsigned long long>(scattered, 0x7fffffffull);
        stats->launches = 2;  // trace kernel + resolve kernel
        cudaEventElapsedTime(&stats->kernel_ms, scene->ev[0], scene->ev[1]);
        cudaEventElapsedTime(&stats->total_ms, scene->ev[0], scene->ev[2]);
    }
    return FT_OK;
}

}  // namespace

extern "C" {

FT_API int ft_abi_version(void) { return FT_ABI_VERSION; }

FT_API const char *ft_last_error(void) { return g_last_error.c_str(); }

FT_API int ft_device_count(int *count) {
    if (!count) return fail(FT_ERR_INVALID_ARGUMENT, "count is null");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    return FT_OK;
}

FT_API float ft_lens_create(float field_of_view) { return ftc::lens_create(field_of_view); }

FT_API int ft_camera_look_at(const float position[3], const float look_at[3], const float up[3], float near_plane_size,
                             ft_camera *out) {
    if (!position || !look_at || !up || !out) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    ftc::camera_look_at(position, look_at, up, near_plane_size, out);
    return FT_OK;
}

FT_API int ft_scene_create(const ft_scene_desc *desc, int device, ft_scene **out) {
    if (!out) return fail(FT_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    ft_scene *s = new (std::nothrow) ft_scene();
    if (!s) return fail(FT_ERR_OUT_OF_MEMORY, "host allocation failed");
    std::string err;
    int st;
    // Spatial lookups are built on the device once they are large (>= 2^24 cell x child pair tests; the host build
    // of the reference's 1024-primitive scenes takes ~10 ms).  FT_GRID_BUILD=gpu|cpu forces either (tests).
    ftc::GridBuildFn grid_fn = nullptr;
    long long gpu_min_pairs = 1ll << 24;
    if (device != FT_DEVICE_NONE) {
        int nd = 0;
        if (cudaGetDeviceCount(&nd) == cudaSuccess && device >= 0 && device < nd) grid_fn = ftc::gpu_grid_build;
    }
    if (const char *gb = getenv("FT_GRID_BUILD")) {
        if (!strcmp(gb, "cpu")) grid_fn = nullptr;
        else if (!strcmp(gb, "gpu")) gpu_min_pairs = 0;
    }
    try {
        st = ftc::compile_scene(desc, &s->host, &err, grid_fn, (void *)(intptr_t)device, gpu_min_pairs);
    } catch (const std::bad_alloc &) {
        delete s;
        return fail(FT_ERR_OUT_OF_MEMORY, "host allocation failed while compiling the scene");
    } catch (const std::exception &ex) {
        delete s;
        return fail(FT_ERR_INVALID_ARGUMENT, std::string("scene compiler: ") + ex.what());
    }
    if (st != FT_OK) {
        delete s;
        return fail(st, err);
    }
    if (device == FT_DEVICE_NONE) {  // compile-only handle: introspection works, every compute call fails
        s->device = FT_DEVICE_NONE;
        *out = s;
        return FT_OK;
    }
    // no CPU fallback: a scene only renders on a device
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete s;
        return fail(FT_ERR_NO_DEVICE, std::string("no CUDA device available (") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                          "); fraytracer_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) {
        delete s;
        return fail(FT_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    }
    s->device = device;
    DeviceGuard guard(device);
    if (!guard.ok) {
        delete s;
        return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    }
    auto bail = [&](cudaError_t ce, const char *what) {
        int code = cuda_fail(ce, what);
        s->release();
        delete s;
        return code;
    };
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    s->sm_count = prop.multiProcessorCount;
    const ftc::CompiledScene &h = s->host;
    int64_t bytes = 0;
    if ((e = upload(h.nodes, &s->dev.nodes, &bytes)) != cudaSuccess) return bail(e, "upload nodes");
    if ((e = upload(h.cells, &s->dev.cells, &bytes)) != cudaSuccess) return bail(e, "upload cells");
    if ((e = upload(h.fat_entries, &s->dev.entries, &bytes)) != cudaSuccess) return bail(e, "upload entries");
    if ((e = upload(h.links, &s->dev.links, &bytes)) != cudaSuccess) return bail(e, "upload links");
    if ((e = upload(h.objects, &s->dev.objects, &bytes)) != cudaSuccess) return bail(e, "upload objects");
    if ((e = upload(h.obj_links, &s->dev.obj_links, &bytes)) != cudaSuccess) return bail(e, "upload obj_links");
    if ((e = upload(h.materials, &s->dev.materials, &bytes)) != cudaSuccess) return bail(e, "upload materials");
    if ((e = upload(h.lights, &s->dev.lights, &bytes)) != cudaSuccess) return bail(e, "upload lights");
    s->device_bytes = bytes;
    s->dev.root_word = h.root_word;
    s->dev.root_object = h.root_object;
    s->dev.n_lights = h.n_lights;
    s->dev.step_cap = 1 << 22;
    s->dev.bg[0] = h.bg[0];
    s->dev.bg[1] = h.bg[1];
    s->dev.bg[2] = h.bg[2];
    for (int k = 0; k < 3; k++) {
        s->dev.pg_lo[k] = h.pg_lo[k];
        s->dev.pg_hi[k] = h.pg_hi[k];
        s->dev.esc_lo[k] = h.esc_lo[k];
        s->dev.esc_hi[k] = h.esc_hi[k];
    }
    s->dev.esc_valid = h.esc_valid;
    s->dev.chain_valid = h.chain_valid;
    s->dev.chain_union = h.chain_union;
    s->dev.chain_n_ops = h.chain_n_ops;
    s->dev.chain_ops = h.chain_ops;
    s->dev.chain_mat_valid = h.chain_mat_valid;
    s->dev.chain_mat_links = h.chain_mat_links;
    if (getenv("FT_NO_CHAIN_MAT")) s->dev.chain_mat_valid = 0;  // experiments / tests of the separate arg-min scan
    if (h.chain_valid) {
        s->dev.chain_g1 = h.nodes[(size_t)h.chain_union + 1];
        s->dev.chain_g2 = h.nodes[(size_t)h.chain_union + 2];
        s->dev.chain_g3 = h.nodes[(size_t)h.chain_union + 3];
    }
    if (getenv("FT_NO_CHAIN")) s->dev.chain_valid = 0;  // experiments / tests of the generic interpreter
    // nearest-candidate sub-grid of the chain union's lookup (device_scene.h): built on the device from the uploaded scene
    s->dev.sub_valid = 0;
    // (its headers address candidate records with 24 bits: scenes with 2^24 or more records -- half a gigabyte of them -- keep the scan)
    if (s->dev.chain_valid && !getenv("FT_NO_SUBGRID") && h.fat_entries.size() / 2 < ((size_t)1 << 24)) {
        uint32_t cell_base;
        std::memcpy(&cell_base, &s->dev.chain_g1.w, sizeof(cell_base));
        for (const ftc::GridHost &g : h.grids) {
            if ((uint32_t)g.cell_base != cell_base) continue;
            SubGridJob job;
            for (int k = 0; k < 3; k++) {
                job.amin[k] = g.aabb_min[k];
                job.cell[k] = g.cell_size[k];
                job.count[k] = g.count[k];
            }
            job.cell_base = cell_base;
            job.max_res = 8;
            job.max_subcells = 1ll << 22;
            job.halo = 1;
            if (const char *hh = getenv("FT_SUBGRID_HALO")) job.halo = std::max(0, std::min(8, atoi(hh)));  // experiments
            if (const char *r = getenv("FT_SUBGRID_RES")) job.max_res = std::max(1, std::min(8, atoi(r)));  // experiments
            std::string sub_err;
            if ((e = build_subgrid(s->dev, job, &s->sub, &sub_err)) != cudaSuccess) {
                g_last_error = sub_err;
                return bail(e, "sub-grid build");
            }
            if (s->sub.hdr) {
                s->dev.sub_hdr = s->sub.hdr;
                {
                    const long long off = (long long)s->sub.halo << s->sub.shift;
                    s->dev.sub_hdr_biased = s->sub.hdr + ((off * s->sub.ny + off) * s->sub.nz + off);
                }
                s->dev.sub_idx = s->sub.idx;
                s->dev.sub_shift = s->sub.shift;
                s->dev.sub_off = s->sub.halo << s->sub.shift;
                s->dev.sub_ny = s->sub.ny;
                s->dev.sub_nz = s->sub.nz;
                s->dev.sub_scale = (float)s->sub.res;
                s->dev.sub_dtc_inv = 1.0f / (256.0f * s->sub.dtc_unit);
                for (int k = 0; k < 3; k++) {
                    s->dev.sub_lo[k] = -(float)s->sub.halo;
                    s->dev.sub_hi[k] = (float)(g.count[k] + s->sub.halo);
                }
                s->dev.sub_valid = 1;
                s->device_bytes += s->sub.bytes;
            }
            break;
        }
    }
    s->dev.sub_debug = 0;
    s->hints_enabled = getenv("FT_NO_TILE_HINT") == nullptr;
    if (const char *d = getenv("FT_HINT_PCIE_MODE")) s->hint_pcie_mode = std::min(std::max(atoi(d), 0), 2);  // experiments only
    if (const char *d = getenv("FT_HINT_PCIE_SHIFT")) s->hint_pcie_shift = std::min(std::max(atoi(d), 0), 4);
    if (const char *d = getenv("FT_HINT_SPLIT")) s->hint_split_pct = std::min(std::max(atoi(d), 1), 1000);
    if (const char *d = getenv("FT_SUB_DEBUG")) s->dev.sub_debug = atoi(d);  // tests: 1 = distrust every certificate
    if (getenv("FT_NO_ESCAPE")) s->dev.esc_valid = 0;  // experiments only
    for (int k = 0; k < 3; k++) {  // the fast region (device_scene.h): covered region of the sub-grid, clipped to the escape box
        s->dev.fast_lo[k] = std::numeric_limits<float>::infinity();
        s->dev.fast_hi[k] = -std::numeric_limits<float>::infinity();
        if (!s->dev.sub_valid || getenv("FT_NO_FAST_REGION")) continue;
        float lo = s->dev.sub_lo[k], hi = s->dev.sub_hi[k];
        if (s->dev.esc_valid) {
            const float *g1 = &s->dev.chain_g1.x, *g2 = &s->dev.chain_g2.x;
            lo = std::max(lo, (s->dev.esc_lo[k] - g1[k]) * g2[k]);
            hi = std::min(hi, (s->dev.esc_hi[k] - g1[k]) * g2[k]);
        }
        s->dev.fast_lo[k] = lo;
        s->dev.fast_hi[k] = hi;
    }
    s->sub_valid_built = s->dev.sub_valid;
    for (int k = 0; k < 3; k++) {
        s->fast_lo_built[k] = s->dev.fast_lo[k];
        s->fast_hi_built[k] = s->dev.fast_hi[k];
    }
    s->apply_sub_debug();
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&s->hint_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (auto &ev : s->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&s->d_stats, kStatWords * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc stats");
    if ((e = cudaMemset(s->d_stats, 0, kStatWords * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMemset stats");
    if ((e = cudaMalloc(&s->d_counters, kCounterRing * 16 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc counters");
    if ((e = cudaMallocHost(&s->h_stats, kStatWords * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMallocHost stats");
    *out = s;
    return FT_OK;
}

FT_API int ft_scene_destroy(ft_scene *scene) {
    if (!scene) return FT_OK;
    if (scene->device != FT_DEVICE_NONE) {
        DeviceGuard guard(scene->device);
        if (scene->stream) cudaStreamSynchronize(scene->stream);
        scene->release();
    }
    delete scene;
    return FT_OK;
}

FT_API int ft_scene_info_get(const ft_scene *scene, ft_scene_info *out) {
    if (!scene || !out) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    const ftc::CompiledScene &h = scene->host;
    out->n_nodes = h.n_nodes;
    out->n_primitives = h.n_primitives;
    out->n_grids = (int32_t)h.grids.size();
    out->max_depth = h.max_depth;
    out->n_cells = 0;
    out->n_entries = 0;
    for (const auto &g : h.grids) {
        out->n_cells += g.n_cells;
        out->n_entries += g.n_entries;
    }
    out->device_bytes = scene->device_bytes;
    out->grids_built_on_gpu = h.grids_built_on_gpu;
    out->device = scene->device;
    out->sm_count = scene->sm_count;
    return FT_OK;
}

FT_API int ft_scene_subgrid_info(const ft_scene *scene, ft_subgrid_info *out) {
    if (!scene || !out) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(out, 0, sizeof(*out));
    out->resolution = scene->sub.hdr ? scene->sub.res : 0;
    out->halo = scene->sub.hdr ? scene->sub.halo : 0;
    out->n_subcells = scene->sub.n_subcells;
    out->n_entries = scene->sub.n_entries;
    out->device_bytes = scene->sub.bytes;
    out->threshold = scene->sub.threshold;
    out->build_ms = scene->sub.build_ms;
    return FT_OK;
}

FT_API int ft_scene_set_debug(ft_scene *scene, int32_t sub_debug) {
    if (!scene) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    std::lock_guard<std::mutex> lock(scene->mu);
    scene->dev.sub_debug = sub_debug & 3;
    scene->apply_sub_debug();
    scene->hints_enabled = (sub_debug & 4) == 0 && getenv("FT_NO_TILE_HINT") == nullptr;  // bit 2: natural tile order, no hints
    scene->ext_count = (sub_debug & 8) != 0;  // bit 3: the extension entry points run their instrumented kernels (eval_flops, march_steps)
    return FT_OK;
}

FT_API int ft_scene_set_step_cap(ft_scene *scene, int32_t max_steps) {
    if (!scene || max_steps < 1) return fail(FT_ERR_INVALID_ARGUMENT, "invalid step cap");
    std::lock_guard<std::mutex> lock(scene->mu);
    scene->dev.step_cap = max_steps;
    return FT_OK;
}

FT_API int ft_scene_grid_count(const ft_scene *scene, int32_t *count) {
    if (!scene || !count) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    *count = (int32_t)scene->host.grids.size();
    return FT_OK;
}

FT_API int ft_scene_grid_info(const ft_scene *scene, int32_t grid, ft_grid_info *out) {
    if (!scene || !out || grid < 0 || grid >= (int32_t)scene->host.grids.size())
        return fail(FT_ERR_INVALID_ARGUMENT, "grid index out of range");
    const ftc::GridHost &g = scene->host.grids[(size_t)grid];
    out->node = g.node;
    for (int k = 0; k < 3; k++) {
        out->count[k] = g.count[k];
        out->aabb_min[k] = g.aabb_min[k];
        out->cell_size[k] = g.cell_size[k];
        out->cell_size_inv[k] = g.cell_size_inv[k];
    }
    out->n_entries = g.n_entries;
    return FT_OK;
}

FT_API int ft_scene_grid_read(const ft_scene *scene, int32_t grid, float *cell_centers, int64_t *cell_begin,
                              float *entry_lower_bound, int32_t *entry_item) {
    if (!scene || grid < 0 || grid >= (int32_t)scene->host.grids.size())
        return fail(FT_ERR_INVALID_ARGUMENT, "grid index out of range");
    const ftc::GridHost &g = scene->host.grids[(size_t)grid];
    std::vector<float4> cells((size_t)g.n_cells);
    const size_t n_dev_entries = scene->host.entries.size();
    std::vector<float4> fat(2 * n_dev_entries);
    if (scene->device == FT_DEVICE_NONE) {  // compile-only handle: the arrays that would have been uploaded
        std::copy(scene->host.cells.begin() + g.cell_base, scene->host.cells.begin() + g.cell_base + g.n_cells, cells.begin());
        std::copy(scene->host.fat_entries.begin(), scene->host.fat_entries.begin() + 2 * n_dev_entries, fat.begin());
    } else {  // read back what the kernels actually see
        DeviceGuard guard(scene->device);
        CU(cudaMemcpy(cells.data(), scene->dev.cells + g.cell_base, sizeof(float4) * (size_t)g.n_cells, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(fat.data(), scene->dev.entries, sizeof(float4) * 2 * n_dev_entries, cudaMemcpyDeviceToHost));
    }
    int64_t off = 0;
    for (int64_t c = 0; c < g.n_cells; c++) {
        if (cell_centers) {
            cell_centers[3 * c] = cells[(size_t)c].x;
            cell_centers[3 * c + 1] = cells[(size_t)c].y;
            cell_centers[3 * c + 2] = cells[(size_t)c].z;
        }
        if (cell_begin) cell_begin[c] = off;
        uint32_t b;
        std::memcpy(&b, &cells[(size_t)c].w, 4);
        for (size_t i = b; i < n_dev_entries; i++) {
            const float lb = fat[2 * i].x;
            int32_t child_i;
            std::memcpy(&child_i, &fat[2 * i + 1].z, 4);
            if (child_i < 0) break;  // sentinel
            if (off >= g.n_entries) return fail(FT_ERR_CUDA, "device grid is longer than the compiled grid");
            if (entry_lower_bound) entry_lower_bound[off] = lb;
            if (entry_item) entry_item[off] = child_i;
            off++;
        }
    }
    if (cell_begin) cell_begin[g.n_cells] = off;
    if (off != g.n_entries) return fail(FT_ERR_CUDA, "device grid does not match the compiled grid");
    return FT_OK;
}

FT_API int ft_strip_layout(int32_t width, int32_t height, int32_t strip_width, int32_t world, int32_t *strips_per_rank,
                           int64_t *floats_per_rank) {
    if (width <= 0 || height <= 0 || strip_width <= 0 || world <= 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid strip layout");
    const int strips = (width + strip_width - 1) / strip_width;
    const int spr = (strips + world - 1) / world;
    if (strips_per_rank) *strips_per_rank = spr;
    if (floats_per_rank) *floats_per_rank = (int64_t)(world == 1 ? width : spr * strip_width) * height * 3;
    return FT_OK;
}

FT_API int ft_render_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                            float length, uint32_t flags, int32_t strip_width, int32_t rank, int32_t world,
                            float *d_out_rgb, int32_t *d_out_leaf, void *stream, ft_render_stats *stats) {
    return ft_render_strips_device(scene, camera, width, height, epsilon, length, flags, strip_width, rank, world, 0, -1,
                                   d_out_rgb, d_out_leaf, stream, stats);
}

FT_API int ft_render_strips_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                                   float length, uint32_t flags, int32_t strip_width, int32_t rank, int32_t world,
                                   int32_t first_local_strip, int32_t n_local_strips, float *d_out_rgb,
                                   int32_t *d_out_leaf, void *stream, ft_render_stats *stats) {
    if (!scene || !camera || !d_out_rgb) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (width <= 0 || height <= 0 || strip_width <= 0 || world <= 0 || rank < 0 || rank >= world)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid image size or strip partition");
    int st = check_frame(width, height);
    if (st != FT_OK) return st;
    st = check_ray_params(epsilon, length);
    if (st != FT_OK) return st;
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    if (scene->host.has_glass)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_GLASS (spectral extension): render it with ft_render_spectral");
    if (scene->host.has_cloud)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_CLOUD (volumetric extension): render it with ft_render_cloud");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaStream_t s = (cudaStream_t)stream;
    const bool count = (flags & FT_FLAG_COUNTERS) != 0;
    RenderParams P = make_params(width, height, epsilon, length, strip_width, rank, world);
    if (n_local_strips >= 0) {  // a range of this rank's local strips
        const int spr = (((width + strip_width - 1) / strip_width) + world - 1) / world;
        if (first_local_strip < 0 || first_local_strip + n_local_strips > spr)
            return fail(FT_ERR_INVALID_ARGUMENT, "local strip range out of bounds");
        if (world == 1 && (first_local_strip + n_local_strips) * strip_width > width)
            P.local_cols = width - first_local_strip * strip_width;  // the plain FColor[X,Y] has no padding columns
        else
            P.local_cols = n_local_strips * strip_width;
        P.xl_begin = first_local_strip * strip_width;
        if (P.local_cols <= 0) {
            if (stats) std::memset(stats, 0, sizeof(*stats));
            return FT_OK;
        }
    }
    if (stats) CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    unsigned long long *wc;
    CU(fresh_counter(scene, s, &wc));
    TileHint *hint = hint_attach(scene, P, 0, s);
    cudaError_t e = FT_DISPATCH((flags & FT_FLAG_FAST_MATH), launch_render, scene->dev, *camera, P, d_out_rgb, d_out_leaf, wc, stats ? scene->d_stats : nullptr, count, scene->sm_count, s, false);
    if (e != cudaSuccess) return cuda_fail(e, "render kernel launch");
    hint_finish(scene, hint, s);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        sum_stats(scene->h_stats, stats);
        long long owned = 0;
        for (int xl = P.xl_begin; xl < P.xl_begin + P.local_cols; xl++) {
            const int strip = xl / strip_width;
            const int x = (strip * world + rank) * strip_width + (xl - strip * strip_width);
            owned += x < width;
        }
        stats->primary_rays = owned * height;
        stats->launches = 1;
    }
    return FT_OK;
}

FT_API int ft_render_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                                   float length, uint32_t flags, int32_t strip_width, int32_t rank, int32_t world,
                                   float *const *d_dst, int32_t n_dst, void *stream, ft_render_stats *stats) {
    if (!scene || !camera || !d_dst) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (width <= 0 || height <= 0 || strip_width <= 0 || world <= 0 || rank < 0 || rank >= world || n_dst < 1 || n_dst > 16)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid image size, strip partition or destination count");
    if (check_frame(width, height) != FT_OK) return FT_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < n_dst; k++)
        if (!d_dst[k]) return fail(FT_ERR_INVALID_ARGUMENT, "null destination framebuffer");
    int st = check_ray_params(epsilon, length);
    if (st != FT_OK) return st;
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    if (scene->host.has_glass || scene->host.has_cloud)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses an extension material: render it with ft_render_spectral / ft_render_cloud");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaStream_t s = (cudaStream_t)stream;
    const bool count = (flags & FT_FLAG_COUNTERS) != 0;
    RenderParams P = make_params(width, height, epsilon, length, strip_width, rank, world);
    if (world == 1) P.local_cols = ((width + strip_width - 1) / strip_width) * strip_width;
    if (stats) CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    unsigned long long *wc;
    CU(fresh_counter(scene, s, &wc));
    bool to_host = false;  // a destination in mapped host memory (render_to_host): its stores cross PCIe
    for (int k = 0; k < n_dst; k++) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, d_dst[k]) == cudaSuccess) to_host = to_host || pa.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    TileHint *hint = hint_attach(scene, P, to_host ? 1 : 0, s, to_host);
    cudaError_t e = FT_DISPATCH((flags & FT_FLAG_FAST_MATH), launch_render_gather, scene->dev, *camera, P, d_dst, n_dst, wc, stats ? scene->d_stats : nullptr, count, scene->sm_count, s);
    if (e != cudaSuccess) return cuda_fail(e, "render-gather kernel launch");
    hint_finish(scene, hint, s);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        sum_stats(scene->h_stats, stats);
        long long owned = 0;
        for (int xl = 0; xl < P.local_cols; xl++) {
            const int strip = xl / strip_width;
            owned += (strip * world + rank) * strip_width + (xl - strip * strip_width) < width;
        }
        stats->primary_rays = owned * height;
        stats->launches = 1;
    }
    return FT_OK;
}

FT_API int ft_device_alloc(int device, int64_t bytes, void **d_ptr) {
    if (!d_ptr || bytes <= 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    CU(cudaMalloc(d_ptr, (size_t)bytes));
    return FT_OK;
}
FT_API int ft_device_free(int device, void *d_ptr) {
    if (!d_ptr) return FT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    CU(cudaFree(d_ptr));
    return FT_OK;
}
FT_API int ft_ipc_export(int device, void *d_ptr, uint8_t handle[FT_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == FT_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    if (!d_ptr || !handle) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, sizeof(h));
    return FT_OK;
}
FT_API int ft_ipc_import(int device, const uint8_t handle[FT_IPC_HANDLE_BYTES], void **d_ptr) {
    if (!d_ptr || !handle) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FT_OK;
}
FT_API int ft_ipc_close(int device, void *d_ptr) {
    if (!d_ptr) return FT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    CU(cudaIpcCloseMemHandle(d_ptr));
    return FT_OK;
}

FT_API int ft_render(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon, float length,
                     uint32_t flags, float *out_rgb, int32_t *out_leaf, ft_render_stats *stats) {
    if (!scene || !camera || !out_rgb) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    int st = check_frame(width, height);
    if (st != FT_OK) return st;
    st = check_ray_params(epsilon, length);
    if (st != FT_OK) return st;
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    if (scene->host.has_glass)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_GLASS (spectral extension): render it with ft_render_spectral");
    if (scene->host.has_cloud)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_CLOUD (volumetric extension): render it with ft_render_cloud");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t npix = (size_t)width * height;
    const size_t rgb_bytes = npix * 3 * sizeof(float), leaf_bytes = out_leaf ? npix * sizeof(int32_t) : 0;
    const bool count = (flags & FT_FLAG_COUNTERS) != 0;
    const bool fast = (flags & FT_FLAG_FAST_MATH) != 0;
    // Zero-copy path: when the caller's buffers are page-locked AND mapped into the device's address space
    // (ft_host_alloc / ft_host_register), the kernel stores each finished pixel straight into host memory -- posted
    // PCIe writes of 96-byte runs (8 pixels of one tile column), ~11 GB/s for a 1080p frame, far below the link's
    // rate -- so the frame needs no device->host copy at all and one launch instead of two bands.
    {
        float *m_rgb = nullptr;
        int32_t *m_leaf = nullptr;
        cudaPointerAttributes pa{};
        bool mapped = getenv("FT_NO_ZEROCOPY") == nullptr;
        if (mapped && cudaPointerGetAttributes(&pa, out_rgb) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            m_rgb = (float *)pa.devicePointer;
        else
            mapped = false;
        if (mapped && out_leaf) {
            if (cudaPointerGetAttributes(&pa, out_leaf) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
                m_leaf = (int32_t *)pa.devicePointer;
            else
                mapped = false;
        }
        cudaGetLastError();  // a pageable pointer makes cudaPointerGetAttributes report an error on old drivers
        if (mapped) {
            cudaStream_t s = scene->stream;
            CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
            CU(cudaEventRecord(scene->ev[0], s));
            RenderParams P = make_params(width, height, epsilon, length, width, 0, 1);
            unsigned long long *wc;
            CU(fresh_counter(scene, s, &wc));
            TileHint *hint = hint_attach(scene, P, 1, s, true);
            cudaError_t e = FT_DISPATCH(fast, launch_render, scene->dev, *camera, P, m_rgb, m_leaf, wc, scene->d_stats, count, scene->sm_count, s, true);
            if (e != cudaSuccess) return cuda_fail(e, "render kernel launch");
            CU(cudaEventRecord(scene->ev[1], s));
            hint_finish(scene, hint, s);
            CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CU(cudaEventRecord(scene->ev[2], s));
            CU(cudaStreamSynchronize(s));  // kernel completion makes its system-scope stores visible to the host
            if (stats) {
                std::memset(stats, 0, sizeof(*stats));
                sum_stats(scene->h_stats, stats);
                stats->primary_rays = (int64_t)npix;
                stats->launches = 1;
                cudaEventElapsedTime(&stats->kernel_ms, scene->ev[0], scene->ev[1]);
                cudaEventElapsedTime(&stats->total_ms, scene->ev[0], scene->ev[2]);
            }
            return FT_OK;
        }
    }
    CU(scene->reserve(rgb_bytes + leaf_bytes));
    float *d_rgb = (float *)scene->d_scratch;
    int32_t *d_leaf = out_leaf ? (int32_t *)((char *)scene->d_scratch + rgb_bytes) : nullptr;
    cudaStream_t s = scene->stream, s2 = scene->copy_stream;
    CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    CU(cudaEventRecord(scene->ev[0], s));
    CU(cudaStreamWaitEvent(s2, scene->ev[0], 0));
    // Two column bands on two streams: the first band's pixels travel to the host while the second band
    // renders.  Measured on B200 (1080p, pinned output): 1 band 2.96 ms, 2 bands 2.79 ms, 6 bands 4.7 ms --
    // every extra band adds a persistent-kernel tail, so two it is.
    int bands = (npix >= (size_t)1 << 19) ? 2 : 1;
    if (const char *eb = getenv("FT_BANDS")) bands = std::max(1, atoi(eb));  // experiments only
    const int band_cols = ((width + bands - 1) / bands + 3) & ~3;
    int launches = 0;
    for (int x0 = 0; x0 < width; x0 += band_cols) {
        cudaStream_t bs = (launches & 1) ? s2 : s;
        const int cols = std::min(band_cols, width - x0);
        RenderParams P = make_params(width, height, epsilon, length, cols, 0, 1);
        P.local_cols = cols;
        P.x_offset = x0;
        float *band_rgb = d_rgb + (size_t)x0 * height * 3;
        int32_t *band_leaf = d_leaf ? d_leaf + (size_t)x0 * height : nullptr;
        unsigned long long *wc;
        CU(fresh_counter(scene, bs, &wc));
        TileHint *hint = hint_attach(scene, P, 0, bs);
        cudaError_t e = FT_DISPATCH(fast, launch_render, scene->dev, *camera, P, band_rgb, band_leaf, wc, scene->d_stats, count, scene->sm_count, bs, false);
        if (e != cudaSuccess) return cuda_fail(e, "render kernel launch");
        launches++;
        CU(cudaEventRecord((launches & 1) ? scene->ev[1] : scene->ev[3], bs));  // end of the last kernel on each stream
        hint_finish(scene, hint, bs);
        CU(cudaMemcpyAsync(out_rgb + (size_t)x0 * height * 3, band_rgb, (size_t)cols * height * 3 * sizeof(float),
                           cudaMemcpyDeviceToHost, bs));
        if (out_leaf)
            CU(cudaMemcpyAsync(out_leaf + (size_t)x0 * height, band_leaf, (size_t)cols * height * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, bs));
    }
    if (launches > 1) {
        CU(cudaEventRecord(scene->ev[2], s2));
        CU(cudaStreamWaitEvent(s, scene->ev[2], 0));
    }
    CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(scene->ev[2], s));
    CU(cudaStreamSynchronize(s));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        sum_stats(scene->h_stats, stats);
        stats->primary_rays = (int64_t)npix;
        stats->launches = launches;
        cudaEventElapsedTime(&stats->kernel_ms, scene->ev[0], scene->ev[1]);
        if (launches > 1) {
            float other = 0.f;
            cudaEventElapsedTime(&other, scene->ev[0], scene->ev[3]);
            stats->kernel_ms = std::max(stats->kernel_ms, other);
        }
        cudaEventElapsedTime(&stats->total_ms, scene->ev[0], scene->ev[2]);
    }
    return FT_OK;
}

FT_API int ft_deinterleave_device(const float *d_gathered, int32_t width, int32_t height, int32_t strip_width,
                                  int32_t world, int32_t channels, float *d_out, void *stream) {
    if (width <= 0 || strip_width <= 0 || world <= 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    const int spr = (((width + strip_width - 1) / strip_width) + world - 1) / world;
    return ft_deinterleave_strips_device(d_gathered, width, height, strip_width, world, channels, 0, spr, d_out, stream);
}

FT_API int ft_deinterleave_strips_device(const float *d_gathered, int32_t width, int32_t height, int32_t strip_width,
                                         int32_t world, int32_t channels, int32_t first_local_strip,
                                         int32_t n_local_strips, float *d_out, void *stream) {
    if (!d_gathered || !d_out || width <= 0 || height <= 0 || strip_width <= 0 || world <= 0 || channels <= 0 ||
        first_local_strip < 0 || n_local_strips < 0)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    cudaError_t e = launch_deinterleave(d_gathered, width, height, strip_width, world, channels, first_local_strip,
                                        n_local_strips, d_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "deinterleave kernel launch");
    return FT_OK;
}

FT_API int ft_trace_rays(ft_scene *scene, int64_t n, const ft_ray *rays, uint32_t flags, float *out_rgb, int32_t *out_leaf,
                         int32_t *out_steps, uint32_t *out_shadow, ft_render_stats *stats) {
    if (!scene || n < 0 || (n > 0 && !rays)) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n == 0) return FT_OK;
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    if (scene->host.has_glass)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_GLASS (spectral extension): render it with ft_render_spectral");
    if (scene->host.has_cloud)
        return fail(FT_ERR_UNSUPPORTED_NODE, "the scene uses FT_MATERIAL_CLOUD (volumetric extension): render it with ft_render_cloud");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t N = (size_t)n;
    const size_t ray_bytes = N * sizeof(ft_ray), rgb_bytes = N * 12, i_bytes = N * 4;
    CU(scene->reserve(ray_bytes + rgb_bytes + 3 * i_bytes));
    char *base = (char *)scene->d_scratch;
    float4 *d_rays = (float4 *)base;
    float *d_rgb = (float *)(base + ray_bytes);
    int32_t *d_leaf = (int32_t *)(base + ray_bytes + rgb_bytes);
    int32_t *d_steps = out_steps ? d_leaf + N : nullptr;  // null: the kernel may use the escape proof
    uint32_t *d_shadow = (uint32_t *)(d_leaf + 2 * N);
    cudaStream_t s = scene->stream;
    const bool count = (flags & FT_FLAG_COUNTERS) != 0;
    CU(cudaMemsetAsync(scene->d_stats, 0, kStatWords * sizeof(unsigned long long), s));
    CU(cudaEventRecord(scene->ev[0], s));
    CU(cudaMemcpyAsync(d_rays, rays, ray_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(scene->ev[3], s));
    unsigned long long *wc;
    CU(fresh_counter(scene, s, &wc));
    cudaError_t e = FT_DISPATCH((flags & FT_FLAG_FAST_MATH), launch_trace_rays, scene->dev, d_rays, n, d_rgb, d_leaf, d_steps, d_shadow, wc, scene->d_stats, count, scene->sm_count, s);
    if (e != cudaSuccess) return cuda_fail(e, "trace_rays kernel launch");
    CU(cudaEventRecord(scene->ev[1], s));
    if (out_rgb) CU(cudaMemcpyAsync(out_rgb, d_rgb, rgb_bytes, cudaMemcpyDeviceToHost, s));
    if (out_leaf) CU(cudaMemcpyAsync(out_leaf, d_leaf, i_bytes, cudaMemcpyDeviceToHost, s));
    if (out_steps) CU(cudaMemcpyAsync(out_steps, d_steps, i_bytes, cudaMemcpyDeviceToHost, s));
    if (out_shadow) CU(cudaMemcpyAsync(out_shadow, d_shadow, i_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(scene->h_stats, scene->d_stats, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(scene->ev[2], s));
    CU(cudaStreamSynchronize(s));
    if (stats) {
        sum_stats(scene->h_stats, stats);
        stats->primary_rays = n;
        stats->launches = 1;
        cudaEventElapsedTime(&stats->kernel_ms, scene->ev[3], scene->ev[1]);
        cudaEventElapsedTime(&stats->total_ms, scene->ev[0], scene->ev[2]);
    }
    return FT_OK;
}

FT_API int ft_distance(ft_scene *scene, int64_t n, const float *positions, uint32_t flags, float *out_distance) {
    if (!scene || n < 0 || (n > 0 && (!positions || !out_distance))) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return FT_OK;
    if (scene->device == FT_DEVICE_NONE)
        return fail(FT_ERR_NO_DEVICE, "compile-only scene handle (FT_DEVICE_NONE): fraytracer_b200 has no CPU fallback");
    std::lock_guard<std::mutex> lock(scene->mu);
    DeviceGuard guard(scene->device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t N = (size_t)n;
    CU(scene->reserve(N * 16));
    float *d_pts = (float *)scene->d_scratch;
    float *d_out = d_pts + 3 * N;
    cudaStream_t s = scene->stream;
    CU(cudaMemcpyAsync(d_pts, positions, N * 12, cudaMemcpyHostToDevice, s));
    cudaError_t e = FT_DISPATCH((flags & FT_FLAG_FAST_MATH), launch_distance, scene->dev, d_pts, n, d_out, s);
    if (e != cudaSuccess) return cuda_fail(e, "distance kernel launch");
    CU(cudaMemcpyAsync(out_distance, d_out, N * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return FT_OK;
}

FT_API int ft_render_spectral(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                              float length, int32_t n_wavelengths, int32_t max_bounces, uint64_t seed, float *out_rgb,
                              ft_render_stats *stats) {
    if (!scene || !out_rgb) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    return spectral_core(scene, camera, width, height, epsilon, length, n_wavelengths, max_bounces, seed, whole_frame(width),
                         scene->stream, out_rgb, stats);
}

FT_API int ft_render_spectral_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                                            float length, int32_t n_wavelengths, int32_t max_bounces, uint64_t seed,
                                            int32_t strip_width, int32_t rank, int32_t world, float *const *d_dst, int32_t n_dst,
                                            void *stream, ft_render_stats *stats) {
    StripTarget T;
    int st = make_target(width, strip_width, rank, world, d_dst, n_dst, &T);
    if (st != FT_OK) return st;
    return spectral_core(scene, camera, width, height, epsilon, length, n_wavelengths, max_bounces, seed, T, (cudaStream_t)stream,
                         nullptr, stats);
}

FT_API int ft_render_cloud(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                           float length, const ft_cloud_params *cloud, int32_t samples_per_pixel, uint64_t seed,
                           float *out_rgb, ft_render_stats *stats) {
    if (!scene || !out_rgb) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    return cloud_core(scene, camera, width, height, epsilon, length, cloud, samples_per_pixel, seed, whole_frame(width),
                      scene->stream, out_rgb, stats);
}

FT_API int ft_render_cloud_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                                         float length, const ft_cloud_params *cloud, int32_t samples_per_pixel, uint64_t seed,
                                         int32_t strip_width, int32_t rank, int32_t world, float *const *d_dst, int32_t n_dst,
                                         void *stream, ft_render_stats *stats) {
    StripTarget T;
    int st = make_target(width, strip_width, rank, world, d_dst, n_dst, &T);
    if (st != FT_OK) return st;
    return cloud_core(scene, camera, width, height, epsilon, length, cloud, samples_per_pixel, seed, T, (cudaStream_t)stream,
                      nullptr, stats);
}

FT_API int ft_image_max_device(const float *d_rgb, int64_t n_floats, float *d_max, void *stream) {
    if (!d_rgb || !d_max || n_floats < 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    cudaError_t e = launch_image_max(d_rgb, n_floats, d_max, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "image max kernel launch");
    return FT_OK;
}

FT_API int ft_tonemap_device(const float *d_rgb, int32_t width, int32_t height, float gamma, const float *d_max,
                             float max_value, uint64_t seed, uint8_t *d_out_bgr, int32_t stride, void *stream) {
    if (!d_rgb || !d_out_bgr || width <= 0 || height <= 0 || stride < 3 * width || !(gamma > 0.0f))
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    cudaError_t e = launch_tonemap(d_rgb, width, height, gamma, d_max, max_value, seed, d_out_bgr, stride, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "tonemap kernel launch");
    return FT_OK;
}

FT_API int ft_image_max_strips_device(const float *d_rgb, int32_t width, int32_t height, int32_t strip_width, int32_t rank,
                                      int32_t world, float *d_max, void *stream) {
    if (!d_rgb || !d_max || width <= 0 || height <= 0 || strip_width <= 0 || world <= 0 || rank < 0 || rank >= world)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    cudaError_t e = launch_image_max(d_rgb, (long long)width * height * 3, d_max, (cudaStream_t)stream, height * 3, strip_width,
                                     rank, world);
    if (e != cudaSuccess) return cuda_fail(e, "image max kernel launch");
    return FT_OK;
}

FT_API int ft_tonemap_strips_device(const float *d_rgb, int32_t width, int32_t height, int32_t strip_width, int32_t rank,
                                    int32_t world, float gamma, const float *d_max, float max_value, uint64_t seed,
                                    uint8_t *d_out_bgr, int32_t stride, void *stream) {
    if (!d_rgb || !d_out_bgr || width <= 0 || height <= 0 || stride < 3 * width || !(gamma > 0.0f) || strip_width <= 0 ||
        world <= 0 || rank < 0 || rank >= world)
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    cudaError_t e = launch_tonemap(d_rgb, width, height, gamma, d_max, max_value, seed, d_out_bgr, stride, (cudaStream_t)stream,
                                   strip_width, rank, world);
    if (e != cudaSuccess) return cuda_fail(e, "tonemap kernel launch");
    return FT_OK;
}

FT_API int ft_to_bitmap(int device, const float *rgb, int32_t width, int32_t height, float gamma, uint64_t seed,
                        uint8_t *out_bgr, int32_t stride) {
    if (!rgb || !out_bgr || width <= 0 || height <= 0 || stride < 3 * width || !(gamma > 0.0f))
        return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
        return fail(FT_ERR_NO_DEVICE, "no CUDA device available; fraytracer_b200 has no CPU fallback");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FT_ERR_NO_DEVICE, "cudaSetDevice failed");
    const size_t n = (size_t)width * height * 3, out_bytes = (size_t)stride * height;
    float *d_rgb = nullptr, *d_max = nullptr;
    uint8_t *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_rgb, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_max, sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, out_bytes);
    if (e == cudaSuccess) e = cudaMemset(d_out, 0, out_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_rgb, rgb, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_image_max(d_rgb, (long long)n, d_max, nullptr);
    if (e == cudaSuccess) e = launch_tonemap(d_rgb, width, height, gamma, d_max, 0.f, seed, d_out, stride, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out_bgr, d_out, out_bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_rgb);
    cudaFree(d_max);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "ft_to_bitmap");
    return FT_OK;
}

FT_API int ft_save_bitmap(const char *path, const uint8_t *bgr, int32_t width, int32_t height, int32_t stride) {
    if (!path || !bgr || width <= 0 || height <= 0 || stride < 3 * width) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    if (write_bmp_file(path, bgr, width, height, stride) != 0) return fail(FT_ERR_INVALID_ARGUMENT, std::string("cannot write ") + path);
    return FT_OK;
}

FT_API int ft_host_register(void *ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return FT_OK;
}
FT_API int ft_host_unregister(void *ptr) {
    if (!ptr) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    CU(cudaHostUnregister(ptr));
    return FT_OK;
}
FT_API int ft_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(FT_ERR_INVALID_ARGUMENT, "invalid argument");
    CU(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return FT_OK;
}
FT_API int ft_host_get_device_pointer(void *host_ptr, void **d_ptr) {
    if (!host_ptr || !d_ptr) return fail(FT_ERR_INVALID_ARGUMENT, "null argument");
    CU(cudaHostGetDevicePointer(d_ptr, host_ptr, 0));
    return FT_OK;
}
FT_API int ft_host_free(void *ptr) {
    if (!ptr) return FT_OK;
    CU(cudaFreeHost(ptr));
    return FT_OK;
}

FT_API int ft_measure_fp32_peak(int device, float *tflops_fma, float *sm_mhz_estimate) {
    cudaError_t e = measure_fp32_peak(device, tflops_fma, sm_mhz_estimate);
    if (e != cudaSuccess) return cuda_fail(e, "measure_fp32_peak");
    return FT_OK;
}

}  // extern "C"
