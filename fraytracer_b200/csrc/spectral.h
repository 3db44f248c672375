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
    // Path queues are filled from both ends: continuation segments expected to run INSIDE glass (transmitted into it, or
    // reflected internally) are appended from index 0 upwards, segments in free space from index q_cap - 1 downwards.  The two
    // classes differ by an order of magnitude in length (a few steps to the far side of a glass body against a march through
    // the scene plus shading with its shadow rays); a 32-path chunk never mixes them, so lockstep lanes finish together.
    // count[0] / count[1] = paths of either class.  split == 0: one class (everything from index 0; experiments).
    const float4 *in_q;  // live paths: {origin, dir.x} {dir.y, dir.z, Length, bits(pixel)} {bits(slot), bits(k), -, -}
    const unsigned *in_count;
    float4 *out_q;
    unsigned *out_count;
    unsigned q_cap;
    int split;
    float *slots;  // radiance of path (glass hit h, wavelength k) at h * nl + k
    unsigned long long *work_counter;
    unsigned long long *stats;
};

// count: the instrumented variant (algorithmic flops and march steps into the stats block)
cudaError_t spectral_launch_primary(const DevScene &S, const SpectralParams &P, bool count, int sm_count, cudaStream_t stream);
cudaError_t spectral_launch_spawn(const DevScene &S, const SpectralParams &P, unsigned n_hits, cudaStream_t stream);
cudaError_t spectral_launch_wave(const DevScene &S, const SpectralParams &P, long long max_items, bool count, int sm_count,
                                 cudaStream_t stream);
cudaError_t spectral_launch_resolve(const SpectralParams &P, unsigned n_hits, cudaStream_t stream);
