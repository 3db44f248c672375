// util_kernels.cu -- strip de-interleave after the multi-GPU all-gather, and the FP32 peak probe.
#include "kernels.h"

// gathered: [world][n_local_strips][strip_width][height][channels], holding local strips
// first_local_strip .. first_local_strip + n_local_strips - 1 of every rank  ->  out: [width][height][channels]
// (x-major FColor[X,Y]).  One thread per SOURCE scalar: a strip column is contiguous on both sides, so reads and
// writes are coalesced.  Padding columns (image column >= width) are dropped.
__global__ void __launch_bounds__(256) deinterleave_kernel(const float *__restrict__ g, int width, int height,
                                                           int strip_width, int world, int channels,
                                                           int first_local_strip, int n_local_strips,
                                                           float *__restrict__ out) {
    const long long col_elems = (long long)height * channels;
    const long long cols_per_rank = (long long)n_local_strips * strip_width;
    const long long total = (long long)world * cols_per_rank * col_elems;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long col = i / col_elems, rem = i - col * col_elems;
        const int rank = (int)(col / cols_per_rank);
        const int lc = (int)(col - (long long)rank * cols_per_rank);
        const int local = lc / strip_width, within = lc - local * strip_width;
        const long long x = ((long long)(first_local_strip + local) * world + rank) * strip_width + within;
        if (x < width) out[x * col_elems + rem] = __ldg(g + i);
    }
}

cudaError_t launch_deinterleave(const float *gathered, int width, int height, int strip_width, int world, int channels,
                                int first_local_strip, int n_local_strips, float *out, cudaStream_t stream) {
    const long long total = (long long)world * n_local_strips * strip_width * height * channels;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    deinterleave_kernel<<<(unsigned)blocks, 256, 0, stream>>>(gathered, width, height, strip_width, world, channels,
                                                              first_local_strip, n_local_strips, out);
    return cudaGetLastError();
}

// Dependent-free FFMA streams: 8 independent accumulators per thread, 8 resident warps per SMSP.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a, float b) {
    float r0 = threadIdx.x, r1 = r0 + 1.f, r2 = r0 + 2.f, r3 = r0 + 3.f, r4 = r0 + 4.f, r5 = r0 + 5.f, r6 = r0 + 6.f, r7 = r0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
            r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
        }
    }
    const float s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 12345.678f) out[0] = s;  // never true; keeps the loop alive
}

cudaError_t measure_fp32_peak(int device, float *tflops_fma, float *sm_mhz_estimate) {
    // the caller's current device is restored on every exit (the host thread may be torch's)
    struct Restore {
        int prev = -1;
        Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
        ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    float *d = nullptr;
    e = cudaMalloc(&d, 256);
    if (e != cudaSuccess) return e;
    cudaEvent_t a = nullptr, b = nullptr;
    if ((e = cudaEventCreate(&a)) != cudaSuccess || (e = cudaEventCreate(&b)) != cudaSuccess) {
        if (a) cudaEventDestroy(a);
        cudaFree(d);
        return e;
    }
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    float best = 0.f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(a);
        ffma_peak_kernel<<<blocks, 256>>>(d, iters, 1.000001f, 1e-7f);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    if (tflops_fma) *tflops_fma = best;
    // 128 FP32 lanes per SM, 2 flops per FFMA
    if (sm_mhz_estimate) *sm_mhz_estimate = (float)(best * 1e12 / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6);
    return cudaSuccess;
}
