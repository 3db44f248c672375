// microbenchmark: do packed f32x2 ops save issue slots next to ALU work?  (development tool)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0 = pk(x0, x1), p1 = pk(x2, x3), p2 = pk(x4, x5), p3 = pk(x6, x7), pa = pk(a, a), pb = pk(b, b);
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {  // 8 scalar FP (4 add + 4 mul) + 4 int
                x0 = __fadd_rn(x0, b); x1 = __fmul_rn(x1, a); x2 = __fadd_rn(x2, b); x3 = __fmul_rn(x3, a);
                x4 = __fadd_rn(x4, b); x5 = __fmul_rn(x5, a); x6 = __fadd_rn(x6, b); x7 = __fmul_rn(x7, a);
            } else {          // same 8 FP operations as 4 packed instructions + 4 int
                p0 = add2(p0, pb); p1 = mul2(p1, pa); p2 = add2(p2, pb); p3 = mul2(p3, pa);
            }
            if (MODE != 2) { i0 = (i0 ^ i1) + 3; i1 = (i1 ^ i2) + 5; i2 = (i2 ^ i3) + 7; i3 = (i3 ^ i0) + 11; }
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(p0 ^ p1 ^ p2 ^ p3) + i0 + i1 + i2 + i3;
    if (s == 1234.5f) out[0] = s;
}
template <int MODE> float run(int iters) {
    float *d; cudaMalloc(&d, 4); cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 4; r++) { cudaEventRecord(a); k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    int iters = 4096;
    printf("scalar 8 FP + 4 INT per group: %.3f ms\n", run<0>(iters));
    printf("packed 4 FP2 + 4 INT per group: %.3f ms\n", run<1>(iters));
    printf("packed 4 FP2 only             : %.3f ms\n", run<2>(iters));
    return 0;
}
