// pcie_store.cu -- microbenchmark: how fast can SM-issued stores fill a page-locked, mapped host framebuffer, as a function of
// the shape of the stores?  (ft_render's zero-copy path writes FColor[X,Y] tiles from inside the render kernel.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_store tools/micro/pcie_store.cu && build/pcie_store
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);      \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// mode 0: linear, every warp writes 512 contiguous bytes (32 x 16 B), warp w -> bytes [512 w, 512 w + 512)
__global__ void k_linear(float4 *out, long long n16) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_float4(1.f, 2.f, 3.f, (float)i);
}

// tile modes: the frame is FColor[X,Y] x-major (column x at x*H*12 bytes); a tile is TW columns x TH rows; `order` maps the
// n-th tile drawn to a tile id (ty fastest).  Each warp draws tiles from a counter like the render kernel does.
// VEC = 16: runs written as 128-bit stores (TH*12/16 lanes per column); VEC = 4: three scalar stores per pixel
template <int TW, int TH, int VEC>
__global__ void k_tiles(float *out, int W, int H, const unsigned *order, unsigned n_tiles, unsigned long long *counter, int spin) {
    const int lane = threadIdx.x & 31;
    const int tiles_y = H / TH;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tiles) break;
        const unsigned tile = order ? order[t] : (unsigned)t;
        const int tx = tile / tiles_y, ty = tile - tx * tiles_y;
        // some fake work so that stores are spread over the kernel like the renderer's
        float acc = (float)lane;
        for (int k = 0; k < spin; k++) acc = acc * 1.0001f + 0.5f;
        if (VEC == 16) {
            constexpr int PER_COL = TH * 12 / 16;  // 128-bit stores per column run
            for (int s = lane; s < TW * PER_COL; s += 32) {
                const int col = s / PER_COL, part = s - col * PER_COL;
                float4 *p = reinterpret_cast<float4 *>(out + 3ll * ((long long)(tx * TW + col) * H + ty * TH) + part * 4);
                *p = make_float4(acc, 2.f, 3.f, 4.f);
            }
        } else {
            for (int s = lane; s < TW * TH; s += 32) {
                const int col = s / TH, row = s - col * TH;
                float *p = out + 3ll * ((long long)(tx * TW + col) * H + ty * TH + row);
                p[0] = acc;
                p[1] = 2.f;
                p[2] = 3.f;
            }
        }
    }
}

// bulk copy variant: the tile's column runs leave shared memory through cp.async.bulk (TMA, no tensor map)
template <int TW, int TH>
__global__ void k_tiles_bulk(float *out, int W, int H, const unsigned *order, unsigned n_tiles, unsigned long long *counter, int spin) {
    __shared__ __align__(128) float sh[4][TW * TH * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_y = H / TH;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tiles) break;
        const unsigned tile = order ? order[t] : (unsigned)t;
        const int tx = tile / tiles_y, ty = tile - tx * tiles_y;
        float acc = (float)lane;
        for (int k = 0; k < spin; k++) acc = acc * 1.0001f + 0.5f;
        // wait until the previous bulk copies have read shared memory
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int s = lane; s < TW * TH * 3; s += 32) sh[warp][s] = acc;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < TW) {
            float *dst = out + 3ll * ((long long)(tx * TW + lane) * H + ty * TH);
            const unsigned src = (unsigned)__cvta_generic_to_shared(&sh[warp][lane * TH * 3]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(TH * 12) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane < TW) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float time_it(const char *name, double bytes, int reps, void (*launch)(void *), void *ctx, cudaStream_t s) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch(ctx);
    CK(cudaStreamSynchronize(s));
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a, s));
        launch(ctx);
        CK(cudaEventRecord(b, s));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    printf("%-58s %8.3f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
    return best;
}

struct Ctx {
    float *dst;
    int W, H;
    const unsigned *order;
    unsigned n_tiles;
    unsigned long long *counter;
    int spin, grid;
    cudaStream_t s;
    float *dev_src;
    size_t bytes;
};

template <int TW, int TH, int VEC>
static void launch_tiles(void *c) {
    Ctx *x = (Ctx *)c;
    CK(cudaMemsetAsync(x->counter, 0, 8, x->s));
    k_tiles<TW, TH, VEC><<<x->grid, 128, 0, x->s>>>(x->dst, x->W, x->H, x->order, x->n_tiles, x->counter, x->spin);
}
template <int TW, int TH>
static void launch_bulk(void *c) {
    Ctx *x = (Ctx *)c;
    CK(cudaMemsetAsync(x->counter, 0, 8, x->s));
    k_tiles_bulk<TW, TH><<<x->grid, 128, 0, x->s>>>(x->dst, x->W, x->H, x->order, x->n_tiles, x->counter, x->spin);
}
static void launch_linear(void *c) {
    Ctx *x = (Ctx *)c;
    k_linear<<<x->grid, 128, 0, x->s>>>((float4 *)x->dst, (long long)(x->bytes / 16));
}
static void launch_memcpy(void *c) {
    Ctx *x = (Ctx *)c;
    CK(cudaMemcpyAsync(x->dst, x->dev_src, x->bytes, cudaMemcpyDeviceToHost, x->s));
}

int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 3840, H = argc > 2 ? atoi(argv[2]) : 2160;
    const size_t bytes = (size_t)W * H * 12;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    float *host, *hdev, *dev;
    CK(cudaHostAlloc(&host, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    CK(cudaHostGetDevicePointer(&hdev, host, 0));
    CK(cudaMalloc(&dev, bytes));
    unsigned long long *counter;
    CK(cudaMalloc(&counter, 8));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("frame %dx%d = %.1f MB, %d SMs\n", W, H, bytes / 1e6, prop.multiProcessorCount);
    Ctx c{hdev, W, H, nullptr, 0, counter, 0, prop.multiProcessorCount * 9, s, dev, bytes};
    time_it("cudaMemcpyAsync D2H (copy engine)", bytes, 5, launch_memcpy, &c, s);
    time_it("linear 512 B per warp, host", bytes, 5, launch_linear, &c, s);
    if (argc > 3 && !strcmp(argv[3], "quick")) {
        // several GPUs at once (one process each, CUDA_VISIBLE_DEVICES): what the host takes in from all of them together.
        // Two seconds of each, so that the processes overlap whatever their start skew.
        for (int pass = 0; pass < 2; pass++) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            const int reps = (int)(2.0 / (bytes / 45e9)) + 1;
            CK(cudaEventRecord(a, s));
            for (int r = 0; r < reps; r++) (pass ? launch_linear : launch_memcpy)(&c);
            CK(cudaEventRecord(b, s));
            CK(cudaEventSynchronize(b));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("sustained %-30s %8.1f GB/s over %.2f s\n", pass ? "linear stores (SMs)" : "cudaMemcpyAsync D2H", (double)bytes * reps / ms / 1e6,
                   ms / 1e3);
        }
        return 0;
    }
    // shuffled tile orders
    auto make_order = [&](int TW, int TH, int kind) {
        const unsigned n = (unsigned)(W / TW) * (H / TH);
        std::vector<unsigned> o(n);
        for (unsigned i = 0; i < n; i++) o[i] = i;
        if (kind == 1) {  // pseudo-random permutation
            unsigned long long st = 88172645463325252ull;
            for (unsigned i = n - 1; i > 0; i--) {
                st ^= st << 13; st ^= st >> 7; st ^= st << 17;
                const unsigned j = (unsigned)(st % (i + 1));
                unsigned t = o[i]; o[i] = o[j]; o[j] = t;
            }
        }
        unsigned *d;
        CK(cudaMalloc(&d, n * 4));
        CK(cudaMemcpy(d, o.data(), n * 4, cudaMemcpyHostToDevice));
        c.order = d;
        c.n_tiles = n;
    };
    for (int spin : {0, 2000}) {
        c.spin = spin;
        printf("-- spin %d (fake work per tile)\n", spin);
        for (int kind = 0; kind < 2; kind++) {
            const char *kn = kind ? "random order" : "natural order";
            char name[128];
            make_order(4, 8, kind);
            snprintf(name, sizeof name, "tile 4x8, 96 B runs, STG.128, %s", kn);
            time_it(name, bytes, 4, launch_tiles<4, 8, 16>, &c, s);
            snprintf(name, sizeof name, "tile 4x8, scalar STG.32, %s", kn);
            time_it(name, bytes, 4, launch_tiles<4, 8, 4>, &c, s);
            snprintf(name, sizeof name, "tile 4x8, cp.async.bulk 96 B, %s", kn);
            time_it(name, bytes, 4, launch_bulk<4, 8>, &c, s);
            make_order(4, 16, kind);
            snprintf(name, sizeof name, "tile 4x16, 192 B runs, STG.128, %s", kn);
            time_it(name, bytes, 4, launch_tiles<4, 16, 16>, &c, s);
            snprintf(name, sizeof name, "tile 4x16, cp.async.bulk 192 B, %s", kn);
            time_it(name, bytes, 4, launch_bulk<4, 16>, &c, s);
            make_order(2, 16, kind);
            snprintf(name, sizeof name, "tile 2x16, 192 B runs, STG.128, %s", kn);
            time_it(name, bytes, 4, launch_tiles<2, 16, 16>, &c, s);
            make_order(1, 32, kind);
            snprintf(name, sizeof name, "tile 1x32, 384 B runs, STG.128, %s", kn);
            time_it(name, bytes, 4, launch_tiles<1, 32, 16>, &c, s);
            make_order(4, 32, kind);
            snprintf(name, sizeof name, "tile 4x32, 384 B runs, STG.128, %s", kn);
            time_it(name, bytes, 4, launch_tiles<4, 32, 16>, &c, s);
            snprintf(name, sizeof name, "tile 4x32, cp.async.bulk 384 B, %s", kn);
            time_it(name, bytes, 4, launch_bulk<4, 32>, &c, s);
        }
    }
    // device destination for contrast
    c.dst = dev;
    c.spin = 0;
    make_order(4, 8, 0);
    time_it("tile 4x8, STG.128, natural order, DEVICE memory", bytes, 4, launch_tiles<4, 8, 16>, &c, s);
    return 0;
}
