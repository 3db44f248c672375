// spectral.h -- parameters and launchers of the spectral / refractive extension (spectral_kernels.cu).
// EXTENSION: no reference implementation exists.
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

struct SpectralParams {
    ft_camera cam;
    int width, height;
    float epsilon, length, max_size;
    int nl;           // wavelengths per pixel
    int max_bounces;  // glass interactions allowed per path
    int wave;         // glass interactions the paths of this launch have behind them
    uint32_t seed_lo, seed_hi;
    long long n_items;  // wave 0: pixel tiles of the rank's compact strip space * 32
    StripTarget T;      // the rank's strips and the destination framebuffers
    float4 *hit_q;      // glass hits of wave 0: {origin, Length} {direction, bits(pixel)} {normal, bits(material)}
    unsigned *hit_count;
    const float4 *in_q;  // live paths: {origin, dir.x} {dir.y, dir.z, Length, bits(pixel)} {bits(slot), bits(k), -, -}
    const unsigned *in_count;
    float4 *out_q;
    unsigned *out_count;
    float *slots;  // radiance of path (glass hit h, wavelength k) at h * nl + k
    unsigned long long *work_counter;
    unsigned long long *stats;
};

cudaError_t spectral_launch_primary(const DevScene &S, const SpectralParams &P, int sm_count, cudaStream_t stream);
cudaError_t spectral_launch_spawn(const DevScene &S, const SpectralParams &P, unsigned n_hits, cudaStream_t stream);
cudaError_t spectral_launch_wave(const DevScene &S, const SpectralParams &P, long long max_items, int sm_count, cudaStream_t stream);
cudaError_t spectral_launch_resolve(const SpectralParams &P, unsigned n_hits, cudaStream_t stream);
