// subgrid.h -- host interface of subgrid_build.cu (the nearest-candidate sub-grid, device_scene.h)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "device_scene.h"

struct SubGridJob {
    float amin[3], cell[3];  // aabbMin and cellSize of the chain union's spatial lookup
    int count[3];
    uint32_t cell_base;      // its first cell in cells[]
    int max_res;             // 8, 4, 2 or 1 sub-cells per axis at most
    int halo;                // cells of halo around the lookup's box that the fine grid covers as well
    long long max_subcells;  // budget: the resolution is halved until cells * res^3 fits
};
struct SubGrid {
    uint2 *hdr = nullptr;  // device arrays (cudaFree by the owner); null when no sub-grid was built
    uint32_t *idx = nullptr;
    int res = 0, shift = 0, halo = 0, ny = 0, nz = 0;
    long long n_subcells = 0, n_entries = 0, bytes = 0;
    float threshold = 0.f, build_ms = 0.f, dtc_unit = 0.f;
};
// `S` must already hold the uploaded nodes / cells / entries.  Returns cudaSuccess with out->hdr == nullptr when the lookup is
// too large to encode (the kernels then always run the reference's scan).
cudaError_t build_subgrid(const DevScene &S, const SubGridJob &job, SubGrid *out, std::string *err);
