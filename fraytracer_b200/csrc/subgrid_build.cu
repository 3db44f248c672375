// subgrid_build.cu -- the nearest-candidate sub-grid of a chain scene's spatial lookup U (device_scene.h "sub-grid",
// trace_device.cuh sub_lookup).  Built on the device right after the compiled scene has been uploaded.
//
// The reference evaluates U(p) by scanning the cell's candidate list (SdfForm.fs:22-34); sub_lookup proves that the value (and
// the arg-min that becomes the material leaf) only depends on the candidates that can be nearest to p.  Here every cell is
// cut into S^3 sub-cells and each sub-cell receives the candidates of ITS CELL'S OWN list with
//
//        D_i(centre) <= min_j D_j(centre) + 2 hd + tol
//
// (D the primitive's Distance, hd the half diagonal of the sub-cell inflated for the rounding of the cell coordinate, tol a
// scale-relative allowance for every float rounding involved): Distance functions are 1-Lipschitz, so a dropped candidate is
// farther than the list's nearest one everywhere in the sub-cell by more than tol.  The list is ordered nearest first, so
// that the running minimum of sub_lookup is tight after the first candidate.  No exact arithmetic is needed here (margins),
// but the kernels reuse the bit-exact device evaluator for the distances.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>
#include <string>

#define FT_NS ft_sub
#define FT_WITH_CONE 1
#include "kernels.h"
#include "subgrid.h"
#include "trace_device.cuh"

namespace ft_sub {

struct SubParams {
    float amin[3], cell[3];  // U's aabbMin and cellSize
    int cx, cy, cz;          // cell counts
    int S, shift;            // sub-cells per cell and axis (1 << shift)
    int halo;                // cells of halo around the box
    int nx, ny, nz;          // fine cells per axis: (count + 2 halo) S
    uint32_t cell_base;      // first cell of U in cells[]
    float thr;               // 2 hd + tol
    float dtc_unit;          // quantum of the header's lower bound of dist(p, cell centre) (half a cell diagonal / 255)
    unsigned n_sub;          // nx ny nz
};

// fine cell t = (gx ny + gy) nz + gz: its centre and the reference cell whose list it draws from -- clamp(floor(c)) per axis,
// the border cell for the halo (SdfBoundary.fs:276-282)
__device__ __forceinline__ V3 sub_centre(const SubParams &P, unsigned t, uint32_t &cell) {
    const int gz = (int)(t % P.nz), gy = (int)((t / P.nz) % P.ny), gx = (int)(t / (P.nz * P.ny));
    const int off = P.halo << P.shift;
    const int fx = gx - off, fy = gy - off, fz = gz - off;  // fine coordinates relative to aabbMin, negative in the halo
    const int ix = min(max(fx >> P.shift, 0), P.cx - 1), iy = min(max(fy >> P.shift, 0), P.cy - 1),
              iz = min(max(fz >> P.shift, 0), P.cz - 1);
    cell = (uint32_t)((ix * P.cy + iy) * P.cz + iz);
    const float inv = 1.0f / (float)P.S;
    return V3{P.amin[0] + (((float)fx + 0.5f) * inv) * P.cell[0], P.amin[1] + (((float)fy + 0.5f) * inv) * P.cell[1],
              P.amin[2] + (((float)fz + 0.5f) * inv) * P.cell[2]};
}

// A lower bound, in quanta of P.dtc_unit (0 .. 255, rounded down and then lowered by one more quantum: the margin of
// sub_lookup's certificate (1)), of dist(p, centre of the reference cell) over all positions p the kernels can assign to fine
// cell t.  The fine cell is inflated by 1e-3 of its size for the rounding of the cell coordinate it is derived from.
__device__ __forceinline__ unsigned sub_dtc_quanta(const SubParams &P, unsigned t) {
    const int gz = (int)(t % P.nz), gy = (int)((t / P.nz) % P.ny), gx = (int)(t / (P.nz * P.ny));
    const int off = P.halo << P.shift;
    const int f[3] = {gx - off, gy - off, gz - off}, cnt[3] = {P.cx, P.cy, P.cz};
    double s2 = 0.0;
    for (int k = 0; k < 3; k++) {
        const int i = min(max(f[k] >> P.shift, 0), cnt[k] - 1);
        const double centre = ((double)i + 0.5) * (double)P.S;  // in fine-cell units from aabbMin
        double gap = fabs(((double)f[k] + 0.5) - centre) - 0.5 - 1e-3;
        gap = gap > 0.0 ? gap * ((double)P.cell[k] / (double)P.S) : 0.0;
        s2 += gap * gap;
    }
    const double q = floor(sqrt(s2) / (double)P.dtc_unit) - 1.0;
    return q <= 0.0 ? 0u : (q >= 255.0 ? 255u : (unsigned)q);
}

// pass 1: nearest candidate distance at the sub-cell centre and the number of candidates within the threshold
__global__ void __launch_bounds__(128) sub_count_kernel(const __grid_constant__ DevScene S, const __grid_constant__ SubParams P,
                                                        float *__restrict__ dmin_out, uint32_t *__restrict__ count) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n_sub) return;
    uint32_t cell;
    const V3 c = sub_centre(P, t, cell);
    const uint32_t first = __float_as_uint(ld4(S.cells + P.cell_base + cell).w);
    Acc acc{0ull, 0ull};
    float dmin = __int_as_float(0x7f800000);
    for (uint32_t k = first;; k++) {
        const Entry e = ld_entry(S, k);
        if (e.lb() == __int_as_float(0x7f800000)) break;  // sentinel
        const float d = prim_distance<false>(S, e.word(), e.bound(), c, acc);
        dmin = d < dmin ? d : dmin;
    }
    const float lim = dmin + P.thr;
    uint32_t n = 0;
    for (uint32_t k = first;; k++) {
        const Entry e = ld_entry(S, k);
        if (e.lb() == __int_as_float(0x7f800000)) break;
        n += !(prim_distance<false>(S, e.word(), e.bound(), c, acc) > lim);  // NaN distances are kept
    }
    dmin_out[t] = dmin;
    // what the index array holds: candidates 1 .. n-1 (the nearest one lives in the header); n > 255 is too long to encode:
    // no list, sub_lookup leaves such sub-cells to the scan
    count[t] = (n > 255u || n == 0u) ? 0u : n - 1u;
}

// pass 2: the list, nearest first (ties by list position)
__global__ void __launch_bounds__(128) sub_fill_kernel(const __grid_constant__ DevScene S, const __grid_constant__ SubParams P,
                                                       const float *__restrict__ dmin_in, const uint32_t *__restrict__ count,
                                                       const uint32_t *__restrict__ offset, uint2 *__restrict__ hdr,
                                                       uint32_t *__restrict__ idx_out) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n_sub) return;
    const uint32_t off = offset[t];
    uint32_t cell;
    const V3 c = sub_centre(P, t, cell);
    const uint32_t first = __float_as_uint(ld4(S.cells + P.cell_base + cell).w);
    Acc acc{0ull, 0ull};
    const float lim = dmin_in[t] + P.thr;
    float key[255];
    uint32_t out[255];
    uint32_t m = 0;
    for (uint32_t k = first; m < 255u; k++) {
        const Entry e = ld_entry(S, k);
        if (e.lb() == __int_as_float(0x7f800000)) break;
        const float d = prim_distance<false>(S, e.word(), e.bound(), c, acc);
        if (d > lim) continue;
        // insertion sort by (d, k); NaN keys go last
        uint32_t j = m;
        while (j > 0 && (key[j - 1] > d || key[j - 1] != key[j - 1])) {
            key[j] = key[j - 1];
            out[j] = out[j - 1];
            j--;
        }
        key[j] = d;
        out[j] = k;
        m++;
    }
    // the count pass and this pass see the same distances, so m - 1 == count[t] whenever the list is encodable; a list the
    // count pass refused (more than 255 candidates) stays empty
    const bool ok = m > 0u && m - 1u == count[t];
    // header: {(first << 8) | count, (dtc quanta << 24) | nearest candidate}; entries[] indices are < 2^24 (checked by the caller)
    hdr[t] = ok ? make_uint2((off << 8) | m, (sub_dtc_quanta(P, t) << 24) | out[0]) : make_uint2(0u, 0u);
    for (uint32_t j = 1; ok && j < m; j++) idx_out[off + j - 1u] = out[j];
}

}  // namespace ft_sub

#define SG_CU(call)                                                                               \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            if (err) *err = std::string("sub-grid build: ") + #call + ": " + cudaGetErrorString(e__); \
            cleanup();                                                                            \
            return e__;                                                                           \
        }                                                                                         \
    } while (0)

cudaError_t build_subgrid(const DevScene &S, const SubGridJob &job, SubGrid *out, std::string *err) {
    *out = SubGrid{};
    float *d_dmin = nullptr;
    uint32_t *d_count = nullptr, *d_off = nullptr, *d_idx = nullptr;
    uint2 *d_hdr = nullptr;
    void *d_tmp = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_dmin); cudaFree(d_count); cudaFree(d_off); cudaFree(d_tmp);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    const int halo = job.halo < 0 ? 0 : job.halo;
    const long long n_cells = (long long)(job.count[0] + 2 * halo) * (job.count[1] + 2 * halo) * (job.count[2] + 2 * halo);
    int res = job.max_res;
    while (res > 1 && n_cells * res * res * res > job.max_subcells) res >>= 1;
    SG_CU(cudaEventCreate(&e0));
    SG_CU(cudaEventCreate(&e1));
    SG_CU(cudaEventRecord(e0, 0));
    for (;; res >>= 1) {
        // half extent of a sub-cell in world space, inflated for the rounding of c = fl(fl(p - aabbMin) * cellSizeInv) (each
        // factor within 2^-23, cellSizeInv * cellSize within 2^-22 of 1): 1e-6 of the lookup's reach per axis is ample
        double hd2 = 0.0, scale = 1.0;
        for (int k = 0; k < 3; k++) {
            const double reach = std::fabs((double)job.amin[k]) + (double)(job.count[k] + halo) * job.cell[k];
            const double half = 0.5 / res * job.cell[k] * (1.0 + 1e-5) + 1e-6 * reach;
            hd2 += half * half;
            scale = std::max(scale, 2.0 * reach);
        }
        const double tol = scale * (1.0 / 4096.0);
        ft_sub::SubParams P;
        for (int k = 0; k < 3; k++) { P.amin[k] = job.amin[k]; P.cell[k] = job.cell[k]; }
        P.cx = job.count[0]; P.cy = job.count[1]; P.cz = job.count[2];
        P.S = res;
        P.shift = res == 8 ? 3 : res == 4 ? 2 : res == 2 ? 1 : 0;
        P.halo = halo;
        P.nx = (job.count[0] + 2 * halo) * res;
        P.ny = (job.count[1] + 2 * halo) * res;
        P.nz = (job.count[2] + 2 * halo) * res;
        P.cell_base = job.cell_base;
        P.thr = (float)(2.0 * std::sqrt(hd2) + tol);
        P.dtc_unit = (float)(0.5 * std::sqrt((double)job.cell[0] * job.cell[0] + (double)job.cell[1] * job.cell[1] +
                                             (double)job.cell[2] * job.cell[2]) / 255.0);
        const long long n_sub = n_cells * res * res * res;
        if (n_sub >= (1ll << 31)) { if (res > 1) continue; cleanup(); return cudaSuccess; }
        P.n_sub = (unsigned)n_sub;
        cudaFree(d_dmin); cudaFree(d_count); cudaFree(d_off); cudaFree(d_tmp);
        d_dmin = nullptr; d_count = d_off = nullptr; d_tmp = nullptr;
        SG_CU(cudaMalloc(&d_dmin, sizeof(float) * (size_t)n_sub));
        SG_CU(cudaMalloc(&d_count, sizeof(uint32_t) * ((size_t)n_sub + 1)));
        SG_CU(cudaMalloc(&d_off, sizeof(uint32_t) * ((size_t)n_sub + 1)));
        SG_CU(cudaMemset(d_count + n_sub, 0, sizeof(uint32_t)));
        const unsigned blocks = (unsigned)((n_sub + 127) / 128);
        ft_sub::sub_count_kernel<<<blocks, 128>>>(S, P, d_dmin, d_count);
        SG_CU(cudaGetLastError());
        size_t tmp_bytes = 0;
        SG_CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count, d_off, (int)(n_sub + 1)));
        SG_CU(cudaMalloc(&d_tmp, tmp_bytes));
        SG_CU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_count, d_off, (int)(n_sub + 1)));
        uint32_t total = 0;
        SG_CU(cudaMemcpy(&total, d_off + n_sub, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (total >= (1u << 24)) {  // offsets are 24-bit: try a coarser sub-grid
            if (res > 1) continue;
            cleanup();
            return cudaSuccess;  // no sub-grid: the scan always runs
        }
        SG_CU(cudaMalloc(&d_hdr, sizeof(uint2) * (size_t)n_sub));
        SG_CU(cudaMalloc(&d_idx, sizeof(uint32_t) * (size_t)std::max(1u, total)));
        ft_sub::sub_fill_kernel<<<blocks, 128>>>(S, P, d_dmin, d_count, d_off, d_hdr, d_idx);
        SG_CU(cudaGetLastError());
        SG_CU(cudaEventRecord(e1, 0));
        SG_CU(cudaEventSynchronize(e1));
        out->hdr = d_hdr;
        out->idx = d_idx;
        out->res = res;
        out->shift = P.shift;
        out->halo = halo;
        out->ny = P.ny;
        out->nz = P.nz;
        out->n_subcells = n_sub;
        out->n_entries = (long long)total + n_sub;  // the nearest candidate of every sub-cell lives in its header
        out->bytes = (long long)sizeof(uint2) * n_sub + (long long)sizeof(uint32_t) * std::max(1u, total);
        out->threshold = P.thr;
        out->dtc_unit = P.dtc_unit;
        cudaEventElapsedTime(&out->build_ms, e0, e1);
        cleanup();
        return cudaSuccess;
    }
}
