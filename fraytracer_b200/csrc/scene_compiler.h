// scene_compiler.h -- host scene compiler: flat ft_scene_desc -> device-ready arrays.
// Mirrors the *constructors* of the reference (SdfForm.fs, SdfObject.fs, SdfLight.fs,
// SdfBoundary.buildSpatialLookup) so that the compiled data equals what the F# closures capture.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fraytracer_b200.h"
#include "device_scene.h"

namespace ftc {

struct GridHost {
    int32_t node;         // compiled node index of the union
    int32_t count[3];
    float aabb_min[3], cell_size[3], cell_size_inv[3];
    int64_t cell_base;    // first cell in CompiledScene::cells
    int64_t n_cells;
    int64_t n_entries;    // without sentinels
    // per-cell views for introspection (without sentinels)
    std::vector<int64_t> begin;        // n_cells+1, into lb/item below
    std::vector<float> lower_bound;
    std::vector<int32_t> item;
};

struct CompiledScene {
    std::vector<float4> nodes;
    std::vector<float4> cells;
    std::vector<uint2> entries;        // host staging: {bits(LowerBound), child word}
    std::vector<int32_t> entry_child;  // host staging: child position, -1 for sentinels
    std::vector<float4> fat_entries;   // what is uploaded: see device_scene.h
    std::vector<uint32_t> links;
    std::vector<int4> objects;
    std::vector<int32_t> obj_links;
    std::vector<float4> materials;
    std::vector<float4> lights;
    std::vector<GridHost> grids;
    uint32_t root_word = 0;
    int32_t root_object = 0;
    int32_t n_lights = 0;
    float bg[3] = {0, 0, 0};
    float pg_lo[3] = {0, 0, 0}, pg_hi[3] = {0, 0, 0};  // box of the largest grid (scheduling hint)
    float esc_lo[3] = {0, 0, 0}, esc_hi[3] = {0, 0, 0};  // escape box (device_scene.h)
    int32_t esc_valid = 0;
    int32_t chain_valid = 0, chain_n_ops = 0;
    int32_t chain_mat_valid = 0, chain_mat_links = 0;  // device_scene.h
    uint32_t chain_union = 0, chain_ops = 0;
    int32_t n_nodes = 0, n_primitives = 0, max_depth = 0;
    int32_t grids_built_on_gpu = 0;
    int32_t has_cone = 0;   // a form is FT_FORM_CONE (extension primitive): selects the kernel build that evaluates it
    int32_t has_cloud = 0;  // a leaf uses FT_MATERIAL_CLOUD (volumetric extension): only ft_render_cloud renders it
    int32_t has_glass = 0;  // a leaf uses FT_MATERIAL_GLASS (spectral extension): only ft_render_spectral renders it
};

// Optional device backend for SdfBoundary.buildSpatialLookup (grid_build.cu): same lists, built on the GPU.
struct GridBuildJob {
    const float *bounds;  // n x {cx, cy, cz, r}
    int n, cx, cy, cz;
    float amin[3], cell[3], half[3], half_len;
};
struct GridKeyItem {
    float key;
    int32_t item;
};
struct GridBuildResult {
    std::vector<float> centers;        // 3 per cell
    std::vector<long long> begin;      // n_cells + 1
    std::vector<GridKeyItem> entries;  // per cell: sorted {LowerBound, child position}
};
typedef int (*GridBuildFn)(void *ctx, const GridBuildJob &job, GridBuildResult *res, std::string *err);
int gpu_grid_build(void *ctx /* (void*)(intptr_t)device */, const GridBuildJob &job, GridBuildResult *res, std::string *err);

// Returns FT_OK or an FT_ERR_* code with a message in `err`.  Grids with at least `gpu_min_pairs` (cells x children)
// pair tests are built by `grid_fn` when it is given.
int compile_scene(const ft_scene_desc *desc, CompiledScene *out, std::string *err, GridBuildFn grid_fn = nullptr,
                  void *grid_ctx = nullptr, long long gpu_min_pairs = 0);

// Lens.create / Camera.lookAt (Camera.fs:10-14, 33-42)
float lens_create(float field_of_view);
void camera_look_at(const float position[3], const float look_at[3], const float up[3], float nps, ft_camera *out);

}  // namespace ftc
