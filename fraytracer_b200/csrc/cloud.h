// cloud.h -- parameters and launcher of the volumetric cloud extension (cloud_kernels.cu).
// EXTENSION: no reference implementation exists.
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

struct CloudParams {
    ft_camera cam;
    int width, height;
    float epsilon, length, max_size;
    int spp;
    uint32_t seed_lo, seed_hi;
    long long n_tiles;  // 4 x 8 pixel tiles of the rank's compact strip space
    float density, frequency, step, fade, bias, gain;  // ft_cloud_params
    const int *perm;          // device, 256
    const double *gradients;  // device, 768
    StripTarget T;            // the rank's strips and the destination framebuffers
    float *partials;          // [local tile][lane][FT_CLOUD_SAMPLE_GROUPS][3] group sums, combined by cloud_resolve_kernel
    unsigned long long *work_counter;
    unsigned long long *stats;
};

// trace + resolve; count: the instrumented variant (algorithmic flops into the stats block)
cudaError_t cloud_launch(const DevScene &S, const CloudParams &P, bool count, int sm_count, cudaStream_t stream);
