// philox.h -- Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the deterministic dither
// (post_kernels.cu) and of the spectral extension's reflect/transmit choice (spectral_kernels.cu).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FT_PHILOX_HD __host__ __device__ inline
#else
#define FT_PHILOX_HD inline
#endif

FT_PHILOX_HD uint32_t ft_mulhilo(uint32_t a, uint32_t b, uint32_t *hi) {
    const unsigned long long p = (unsigned long long)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

FT_PHILOX_HD void ft_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = ft_mulhilo(0xD2511F53u, c[0], &hi0), lo1 = ft_mulhilo(0xCD9E8D57u, c[2], &hi1);
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
