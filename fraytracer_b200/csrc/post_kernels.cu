// post_kernels.cu -- Image.toColors + Image.toBitmap on the device.
//
//   Image.toColors gamma rng image   (Image.fs:37-50):  max = Max(0.01, max over pixels of max(R,G,B));
//       per pixel  fcolor / max |> gammaInverse (1/gamma) |> toColor rng
//   FColor.gammaInverse              (FColor.fs:50-55):  MathF.Pow per channel
//   FColor.toColor                   (FColor.fs:43-48):  c * 254.5f + rng.range_01() |> MathF.Round |> int |> min 255
//   Image.toBitmap                   (Image.fs:61-72):   24-bit B,G,R; the buffer is built y-flipped and then
//       Array.rev'ed, which undoes the flip and mirrors x: scanline r, column c holds image[X-1-c, r].
//
// The reference draws the dither from ONE System.Random shared by a parallel map (Image.fs:45-50) -- a data race,
// its bitmap is not reproducible run to run.  Here the three draws of pixel (x,y) are Philox4x32-10(key = seed,
// counter = x*Y + y): deterministic, order independent, and uniform in [0,1) like range_01.  MathF.Pow is the C
// runtime's powf in the reference; CUDA's powf differs from it by a few ulp, which can move a byte by one count
// when c*254.5 + dither lands within ~1e-5 of a rounding boundary (tests bound that fraction).
#include <cstdint>
#include <cstdio>

#include "kernels.h"
#include "philox.h"

namespace {

__device__ inline float atomic_max_float(float *addr, float v) {  // v >= 0 in practice; handles any non-NaN
    int *ia = (int *)addr;
    int old = *ia;
    while (v > __int_as_float(old)) {
        const int assumed = old;
        old = atomicCAS(ia, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
    return __int_as_float(old);
}

// strip_width > 0: only the columns of the strips `rank` owns (strip s -> rank s % world) take part -- the rank's share of
// Array2D.maxWith (Image.fs:40-43) when a frame is cut over several GPUs; the shares meet in an all-reduce(MAX)
__global__ void __launch_bounds__(256) image_max_kernel(const float *__restrict__ rgb, long long n, float *out, int column_floats,
                                                        int strip_width, int rank, int world) {
    float m = -__int_as_float(0x7f800000);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (strip_width > 0 && (int)((i / column_floats) / strip_width) % world != rank) continue;
        m = fmaxf(m, __ldg(rgb + i));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = s[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
        if (threadIdx.x == 0) atomic_max_float(out, m);
    }
}

__device__ inline unsigned char to_byte(float c, float inv_gamma, float mx, uint32_t bits) {
    const float v = powf(c / mx, inv_gamma);                  // fcolor / max |> gammaInverse
    const float u = (float)((double)bits * (1.0 / 4294967296.0));  // range_01: float32 of a double in [0,1)
    const float r = rintf(v * 254.5f + u);                    // MathF.Round: half to even
    int k = (int)r;
    k = k < 255 ? k : 255;                                    // |> min 255
    return (unsigned char)(k < 0 ? 0 : k);
}

// 32 x 32 pixel tiles through shared memory: reads run along y (contiguous in FColor[X,Y]), writes along the
// scanline (contiguous in the bitmap).
__global__ void __launch_bounds__(256) tonemap_kernel(const float *__restrict__ rgb, int X, int Y, float inv_gamma,
                                                      const float *__restrict__ max_ptr, float max_value,
                                                      uint32_t k0, uint32_t k1, unsigned char *__restrict__ out, int stride,
                                                      int strip_width, int rank, int world) {
    __shared__ unsigned char tile[32][32 * 3 + 4];
    const float raw_max = max_ptr ? *max_ptr : max_value;
    const float mx = fmaxf(0.01f, raw_max);  // |> MathF.max 0.01f
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int k = threadIdx.x; k < 32 * 32; k += 256) {
        const int tx = k >> 5, ty = k & 31;  // ty fastest: consecutive threads read consecutive y
        const int x = x0 + tx, y = y0 + ty;
        if (x < X && y < Y && (strip_width <= 0 || (x / strip_width) % world == rank)) {
            const long long idx = (long long)x * Y + y;
            uint32_t c[4] = {(uint32_t)idx, (uint32_t)((unsigned long long)idx >> 32), 0u, 0u};
            ft_philox4x32_10(c, k0, k1);
            const float r = __ldg(rgb + 3 * idx), g = __ldg(rgb + 3 * idx + 1), b = __ldg(rgb + 3 * idx + 2);
            tile[ty][3 * tx + 2] = to_byte(r, inv_gamma, mx, c[0]);  // draws in R, G, B order (FColor.fs:44-46)
            tile[ty][3 * tx + 1] = to_byte(g, inv_gamma, mx, c[1]);
            tile[ty][3 * tx + 0] = to_byte(b, inv_gamma, mx, c[2]);  // BitmapColor = {B; G; R} (Image.fs:54-59)
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 32 * 32; k += 256) {
        const int ty = k >> 5, tc = k & 31;  // tc fastest: consecutive threads write consecutive columns
        const int y = y0 + ty, x = x0 + (31 - tc);  // mirrored within the tile
        if (x < X && y < Y && (strip_width <= 0 || (x / strip_width) % world == rank)) {  // other ranks' columns stay untouched
            const int col = X - 1 - x;  // scanline r = y, column c = X-1-x
            unsigned char *p = out + (long long)y * stride + 3ll * col;
            p[0] = tile[ty][3 * (31 - tc) + 0];
            p[1] = tile[ty][3 * (31 - tc) + 1];
            p[2] = tile[ty][3 * (31 - tc) + 2];
        }
    }
}

}  // namespace

cudaError_t launch_image_max(const float *rgb, long long n_floats, float *d_max, cudaStream_t stream, int column_floats,
                             int strip_width, int rank, int world) {
    const float ninf = -__builtin_huge_valf();
    cudaError_t e = cudaMemcpyAsync(d_max, &ninf, sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess || n_floats <= 0) return e;
    long long blocks = (n_floats + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    image_max_kernel<<<(unsigned)blocks, 256, 0, stream>>>(rgb, n_floats, d_max, column_floats > 0 ? column_floats : 1, strip_width, rank,
                                                           world > 0 ? world : 1);
    return cudaGetLastError();
}

cudaError_t launch_tonemap(const float *rgb, int X, int Y, float gamma, const float *d_max, float max_value,
                           unsigned long long seed, unsigned char *out, int stride, cudaStream_t stream, int strip_width, int rank,
                           int world) {
    if (X <= 0 || Y <= 0) return cudaSuccess;
    dim3 grid((unsigned)((X + 31) / 32), (unsigned)((Y + 31) / 32));
    tonemap_kernel<<<grid, 256, 0, stream>>>(rgb, X, Y, 1.0f / gamma, d_max, max_value, (uint32_t)seed,
                                             (uint32_t)(seed >> 32), out, stride, strip_width, rank, world > 0 ? world : 1);
    return cudaGetLastError();
}

// Image.saveBitmap (Image.fs:87-90): 24-bit BMP; `bgr` are top-down scanlines, BMP stores them bottom-up
int write_bmp_file(const char *path, const unsigned char *bgr, int width, int height, int stride) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    const uint32_t row = ((uint32_t)width * 3u + 3u) & ~3u, data = row * (uint32_t)height, off = 54u;
    unsigned char h[54] = {0};
    h[0] = 'B'; h[1] = 'M';
    auto put32 = [&](int at, uint32_t v) { h[at] = (unsigned char)v; h[at + 1] = (unsigned char)(v >> 8); h[at + 2] = (unsigned char)(v >> 16); h[at + 3] = (unsigned char)(v >> 24); };
    put32(2, off + data); put32(10, off); put32(14, 40); put32(18, (uint32_t)width); put32(22, (uint32_t)height);
    h[26] = 1; h[28] = 24; put32(34, data); put32(38, 3780); put32(42, 3780);
    bool ok = fwrite(h, 1, 54, f) == 54;
    const unsigned char pad[4] = {0, 0, 0, 0};
    for (int r = height - 1; ok && r >= 0; r--) {
        ok = fwrite(bgr + (size_t)r * stride, 1, (size_t)width * 3, f) == (size_t)width * 3;
        if (ok && row > (uint32_t)width * 3u) ok = fwrite(pad, 1, row - (uint32_t)width * 3u, f) == row - (uint32_t)width * 3u;
    }
    ok = fclose(f) == 0 && ok;
    return ok ? 0 : 1;
}
