/*
 * fraytracer_b200.h -- C ABI of the B200-native per-ray hot path of FrayTracer.
 *
 * This is the drop-in boundary.  It replaces, with one
 * shared library `libfraytracer_b200.so`, the reference call
 *
 *     scene |> SdfScene.trace |> Image.render epsilon 30f imageSize camera
 *         (src/FrayTracer.Console/Program.fs:90-93,
 *          SdfScene.trace  src/FrayTracer/SdfScene.fs:7-28,
 *          Image.render    src/FrayTracer/Image.fs:26-35)
 *
 * The reference's extension surface is records of opaque F# closures
 * (src/FrayTracer/Types.fs:40-79).  Closures cannot cross an FFI, so the
 * boundary takes the *reified* scene: one flat, blittable description whose
 * entries mirror the reference constructors one to one.  The library compiles
 * the description (grid build included) into device buffers and renders with
 * hand-written sm_100a kernels.  There is no CPU fallback: every entry point
 * that computes returns FT_ERR_NO_DEVICE / FT_ERR_CUDA when no B200 is usable.
 *
 * Conventions
 *   - every function returns an int status (FT_OK == 0); no exception or abort
 *     crosses the ABI; ft_last_error() gives a thread-local message;
 *   - all structs are plain, 4-byte aligned, little-endian, no padding;
 *   - float means IEEE binary32 everywhere (the reference computes in float32);
 *   - ft_scene handles are immutable after creation (README.md:8 "Immutable
 *     scenes") and may be used from any host thread; calls on one handle are
 *     serialised internally.
 */
#ifndef FRAYTRACER_B200_H
#define FRAYTRACER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FT_API __declspec(dllexport)
#else
#define FT_API __attribute__((visibility("default")))
#endif

#define FT_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
#define FT_OK 0
#define FT_ERR_INVALID_ARGUMENT 1
/* reference: failwith "No SdfObjects given." (SdfForm.fs:16,53; SdfObject.fs:14)
 * and failwithf "blub" (SdfForm.fs:71) on an empty composite                */
#define FT_ERR_EMPTY_COMPOSITE 2
/* a node kind the device code cannot evaluate (user-written closures)       */
#define FT_ERR_UNSUPPORTED_NODE 3
#define FT_ERR_CUDA 4
/* composite nesting deeper than FT_MAX_DEPTH                                */
#define FT_ERR_DEPTH 5
#define FT_ERR_NO_DEVICE 6
#define FT_ERR_OUT_OF_MEMORY 7

/* maximum nesting depth of composite forms the device evaluator supports    */
#define FT_MAX_DEPTH 12

/* ---- forms: mirror SdfForm constructors ---------------------------------- */
enum {
    /* SdfForm.Primitive.sphere  (SdfForm.fs:118-135) p = Center.xyz, Radius  */
    FT_FORM_SPHERE = 1,
    /* SdfForm.Primitive.capsule (SdfForm.fs:137-170) p = From.xyz, To.xyz, Radius */
    FT_FORM_CAPSULE = 2,
    /* SdfForm.Primitive.torus   (SdfForm.fs:172-203) p = Center.xyz, Normal.xyz
     * (un-normalised, as the caller gives it), MajorRadius, MinorRadius      */
    FT_FORM_TORUS = 3,
    /* SdfForm.Primitive.triangle (SdfForm.fs:205-268) p = V1.xyz, V2.xyz, V3.xyz, Radius */
    FT_FORM_TRIANGLE = 4,
    /* EXTENSION, not in the reference (BASELINE.json names cones; the reference has none):
     * capped cone, p = From.xyz, To.xyz, FromRadius, ToRadius (exact distance)              */
    FT_FORM_CONE = 5,
    /* SdfForm.union       (SdfForm.fs:14-40)   children = forms              */
    FT_FORM_UNION = 16,
    /* SdfForm.subtract    (SdfForm.fs:42-49)   children = [a, b]             */
    FT_FORM_SUBTRACT = 17,
    /* SdfForm.intersect   (SdfForm.fs:51-67)   children = forms              */
    FT_FORM_INTERSECT = 18,
    /* SdfForm.unionSmooth (SdfForm.fs:69-91)   p[0] = strength, children = forms */
    FT_FORM_UNION_SMOOTH = 19
};

typedef struct ft_form_desc {
    int32_t kind;        /* FT_FORM_*                                          */
    int32_t first_child; /* composites: index into ft_scene_desc.children      */
    int32_t child_count; /* composites: number of child form ids               */
    float p[10];         /* parameters, see the enum                           */
} ft_form_desc;

/* ---- materials: mirror SdfMaterial (SdfMaterial.fs:4-10) ---------------- */
enum {
    FT_MATERIAL_SOLID = 1, /* SdfMaterial.createSolid color                    */
    /* EXTENSION (no reference implementation): dispersive glass, rgb = {A, B, unused} of the Cauchy
     * index n = A + B / lambda^2 (lambda in micrometres).  Only ft_render_spectral accepts scenes that use it.     */
    FT_MATERIAL_GLASS = 2,
    /* EXTENSION (no reference implementation): participating medium bounded by the object's form,
     * rgb = single-scattering albedo.  Only ft_render_cloud accepts scenes that use it.                            */
    FT_MATERIAL_CLOUD = 3
};

typedef struct ft_material_desc {
    int32_t kind;
    float rgb[3];
} ft_material_desc;

/* ---- objects: mirror SdfObject constructors (SdfObject.fs:6-64) ---------- */
enum {
    /* SdfObject.create material form: form = form id, material = material id  */
    FT_OBJECT_LEAF = 1,
    /* SdfObject.union objects: children = object ids                          */
    FT_OBJECT_UNION = 2,
    /* SdfObject.subtract object form: children = [object id], form = form id  */
    FT_OBJECT_SUBTRACT = 3,
    /* SdfObject.intersect object forms: children = [object id, form ids ...]  */
    FT_OBJECT_INTERSECT = 4
};

typedef struct ft_object_desc {
    int32_t kind;
    int32_t form;        /* LEAF: its form; SUBTRACT: the subtracted form      */
    int32_t material;    /* LEAF only                                          */
    int32_t first_child; /* index into ft_scene_desc.children                  */
    int32_t child_count;
} ft_object_desc;

/* ---- lights: mirror SdfLight constructors (SdfLight.fs:6-42) ------------ */
enum {
    FT_LIGHT_DIRECTIONAL = 1, /* SdfLight.directional direction color: v = direction */
    FT_LIGHT_POINT = 2        /* SdfLight.point position color:        v = position  */
};

#define FT_MAX_LIGHTS 32

typedef struct ft_light_desc {
    int32_t kind;
    float v[3];
    float rgb[3];
} ft_light_desc;

/* ---- scene: mirrors the SdfScene record (Types.fs:74-79) ---------------- */
typedef struct ft_scene_desc {
    const ft_form_desc *forms;
    const ft_object_desc *objects;
    const int32_t *children;
    const ft_material_desc *materials;
    const ft_light_desc *lights;
    int32_t n_forms;
    int32_t n_objects;
    int32_t n_children;
    int32_t n_materials;
    int32_t n_lights;
    int32_t root_object;  /* SdfScene.Object                                   */
    float background[3];  /* SdfScene.BackgroundColor                          */
} ft_scene_desc;

/* ---- camera: mirrors the Camera record (Camera.fs:16-22) ---------------- */
typedef struct ft_camera {
    float position[3];
    float forward[3];
    float up_scaled[3];
    float right_scaled[3];
} ft_camera;

/* ---- ray: mirrors the Ray record (Types.fs:9-17), 32 bytes --------------- */
typedef struct ft_ray {
    float origin[3];
    float direction[3];
    float length;
    float epsilon;
} ft_ray;

typedef struct ft_scene ft_scene; /* opaque */

/* what the scene compiler produced, for DESIGN/bench reporting */
typedef struct ft_scene_info {
    int32_t n_nodes;         /* compiled form nodes (given + synthesised)      */
    int32_t n_primitives;
    int32_t n_grids;         /* grid-accelerated unions                        */
    int32_t max_depth;
    int64_t n_cells;         /* over all grids                                 */
    int64_t n_entries;       /* candidate-list entries over all grids          */
    int64_t device_bytes;    /* resident scene bytes in HBM                    */
    int32_t device;          /* CUDA ordinal                                   */
    int32_t sm_count;
    int32_t grids_built_on_gpu; /* spatial lookups built by the device backend  */
    int32_t reserved;
} ft_scene_info;

/* the nearest-candidate sub-grid the device path builds over the spatial lookup of
 * a chain-shaped scene (DESIGN.md "sub-grid"): an acceleration structure of this
 * implementation only -- the reference has no counterpart; results are certified
 * equal to the reference's sequential candidate scan (SdfForm.fs:22-34)        */
typedef struct ft_subgrid_info {
    int32_t resolution;      /* sub-cells per cell and axis (8, 4, 2, 1); 0 = none */
    int32_t halo;            /* cells around the lookup's box the sub-grid covers too */
    int64_t n_subcells;
    int64_t n_entries;       /* candidate references over all sub-cells        */
    int64_t device_bytes;
    float threshold;         /* 2 x half diagonal + tolerance (world units)    */
    float build_ms;          /* device time of the build                       */
} ft_subgrid_info;

/* per-grid introspection (tests compare this against the oracle's grid) */
typedef struct ft_grid_info {
    int32_t node;            /* compiled node id of the union                  */
    int32_t count[3];        /* countX, countY, countZ (SdfBoundary.fs:237-239)*/
    float aabb_min[3];
    float cell_size[3];
    float cell_size_inv[3];
    int64_t n_entries;
} ft_grid_info;

/* counters filled by the render / trace calls */
typedef struct ft_render_stats {
    int64_t primary_rays;
    int64_t primary_hits;
    int64_t shadow_rays;     /* one per light per lit hit (SdfScene.fs:17-19)  */
    int64_t capped_rays;     /* rays stopped by the watchdog step cap          */
    int64_t march_steps;     /* Distance evaluations inside march loops (0 unless FT_FLAG_COUNTERS) */
    int64_t eval_flops;      /* algorithmic f32 operations (0 unless FT_FLAG_COUNTERS) */
    int64_t escaped_rays;    /* rays (primary or shadow) ended by the escape proof: their remaining
                              * half-line stays outside the scene's bounding box, so they cannot hit */
    float kernel_ms;         /* device time of the kernels of this call        */
    float total_ms;          /* device time incl. copies of this call          */
    int32_t launches;        /* kernels launched by this call                  */
    int32_t reserved;
} ft_render_stats;

/* flags for the compute calls */
#define FT_FLAG_DEFAULT 0u
/* collect march_steps / eval_flops with an instrumented kernel variant      */
#define FT_FLAG_COUNTERS 1u
/* contracted (FMA) arithmetic: faster, no longer bit-identical to the
 * reference's individually rounded float32 operations                       */
#define FT_FLAG_FAST_MATH 2u

/* ---- library ------------------------------------------------------------ */
FT_API int ft_abi_version(void);
FT_API const char *ft_last_error(void);
FT_API int ft_device_count(int *count);

/* Lens.create fieldOfView (Camera.fs:10-14): sin(fov * 0.5f), the F# `sin` on
 * float32 evaluates in double and rounds once.                               */
FT_API float ft_lens_create(float field_of_view);
/* Camera.lookAt (Camera.fs:33-42) */
FT_API int ft_camera_look_at(const float position[3], const float look_at[3], const float up[3],
                             float near_plane_size, ft_camera *out);

/* ---- scene compiler ------------------------------------------------------ */
/* Validates and compiles the description: primitive precomputation
 * (SdfForm.fs:148-149,182-188,216-225), boundaries (SdfBoundary.fs:7-54), the
 * uniform grid of every union (SdfBoundary.buildSpatialLookup,
 * SdfBoundary.fs:225-282), upload to `device`.                              */
FT_API int ft_scene_create(const ft_scene_desc *desc, int device, ft_scene **out);
/* device == FT_DEVICE_NONE compiles without uploading: the handle serves ft_scene_info_get and
 * ft_scene_grid_* only (scene validation, tests of the compiler); every compute entry point
 * returns FT_ERR_NO_DEVICE on it.                                                            */
#define FT_DEVICE_NONE (-1)
FT_API int ft_scene_destroy(ft_scene *scene);
FT_API int ft_scene_info_get(const ft_scene *scene, ft_scene_info *out);
/* watchdog: the reference's march loop has no step cap (SdfForm.fs:93-104);
 * rays that exceed `max_steps` Distance evaluations are stopped as misses and
 * counted in ft_render_stats.capped_rays.  Default 1<<22.                   */
FT_API int ft_scene_set_step_cap(ft_scene *scene, int32_t max_steps);

FT_API int ft_scene_subgrid_info(const ft_scene *scene, ft_subgrid_info *out);
/* test / measurement hook: bit 0 of `sub_debug` makes the device path distrust
 * every sub-grid certificate: the sub-grid is detached from the kernel parameters
 * and every Distance evaluation of this handle runs the reference's sequential
 * candidate scan (SdfForm.fs:27-33) until the bit is cleared; bit 2 draws tiles in
 * natural order (no scheduling hints); results must not change.  Bit 3 makes the
 * extension entry points (ft_render_spectral, ft_render_cloud) run their
 * instrumented kernels: ft_render_stats.eval_flops / march_steps are filled     */
FT_API int ft_scene_set_debug(ft_scene *scene, int32_t sub_debug);

FT_API int ft_scene_grid_count(const ft_scene *scene, int32_t *count);
FT_API int ft_scene_grid_info(const ft_scene *scene, int32_t grid, ft_grid_info *out);
/* copies the grid back FROM THE DEVICE: cell_centers[3*ncells] (x-major:
 * cell (x,y,z) at (x*countY + y)*countZ + z), cell_begin[ncells+1],
 * entry_lower_bound[n_entries], entry_item[n_entries] (child position inside
 * the union's child list).  Any pointer may be NULL.                        */
FT_API int ft_scene_grid_read(const ft_scene *scene, int32_t grid, float *cell_centers,
                              int64_t *cell_begin, float *entry_lower_bound, int32_t *entry_item);

/* ---- the hot path -------------------------------------------------------- */
/* Image.render epsilon length {X=width;Y=height} camera (SdfScene.trace scene)
 * (Image.fs:26-35).  out_rgb is the reference's FColor[X,Y]: element (x,y) at
 * (x*height + y), 3 floats each.  out_leaf (optional) receives the index of
 * the leaf SdfObject whose material coloured the pixel (the arg-min of
 * SdfObject.fs:27-46), -1 for background.  Both are HOST pointers.  When they
 * are page-locked and device-mapped (ft_host_register on the pinned FColor[,],
 * or ft_host_alloc) the kernel stores every finished pixel straight into them
 * -- one launch, no staging buffer, no device->host copy; otherwise the frame
 * is rendered in two column bands whose copies overlap the rendering.
 * Scheduling memory: the handle remembers, per frame shape (size, strip
 * partition, kind of destination), what every 4x8-pixel tile cost, and from the
 * third frame of a shape on the tiles are drawn in an order ranked from those
 * costs (most expensive first).  Pixels do not depend on it; it applies to
 * ft_render, ft_render_device, ft_render_strips_device and
 * ft_render_gather_device alike (FT_NO_TILE_HINT in the environment, or bit 2
 * of ft_scene_set_debug, turns it off).                                       */
FT_API int ft_render(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height,
                     float epsilon, float length, uint32_t flags, float *out_rgb, int32_t *out_leaf,
                     ft_render_stats *stats);

/* Same, DEVICE pointers, asynchronous on `stream` (a cudaStream_t; NULL = the
 * legacy default stream).  Renders only the column strips owned by `rank` of
 * `world`: strip s covers columns [s*strip_width, (s+1)*strip_width) and
 * belongs to rank s % world.  The strips land compacted, in ownership order,
 * in out_rgb ([local_strip][strip_width][height][3]) so that N ranks can
 * all-gather equal-sized buffers; world == 1 gives the plain FColor[X,Y].
 * stats (optional, host) is filled only when FT_FLAG_COUNTERS forces a
 * synchronisation.  A scene handle owns ONE block of device counters: a call
 * that passes `stats` must not overlap other launches of the same handle on
 * other streams (their counts would mix); calls with stats == NULL hand the
 * kernel no counter block and may overlap freely.                            */
FT_API int ft_render_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height,
                            float epsilon, float length, uint32_t flags, int32_t strip_width,
                            int32_t rank, int32_t world, float *d_out_rgb, int32_t *d_out_leaf,
                            void *stream, ft_render_stats *stats);

/* Same for a RANGE of the rank's local strips [first_local_strip, first_local_strip + n_local_strips):
 * d_out_rgb / d_out_leaf point at the first column of that range (a chunk of the compact buffer).  Lets a rank
 * render in chunks so that the all-gather of chunk k overlaps the rendering of chunk k+1.                  */
FT_API int ft_render_strips_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height,
                                   float epsilon, float length, uint32_t flags, int32_t strip_width, int32_t rank,
                                   int32_t world, int32_t first_local_strip, int32_t n_local_strips,
                                   float *d_out_rgb, int32_t *d_out_leaf, void *stream, ft_render_stats *stats);

/* Fused render + framebuffer gather (multi-GPU): renders the strips `rank` owns and stores every finished pixel at its
 * final position of the x-major FColor[X,Y] in EACH of the n_dst destination framebuffers (DEVICE pointers: the rank's
 * own buffer and its peers', opened with ft_ipc_import) -- peer stores over NVLink / NVSwitch from inside the render
 * kernel, replacing compact strips + NCCL all-gather + de-interleave.  Pixels of other ranks' strips are not touched.
 * The caller orders frames: after the call, a collective on the same stream (a 4-byte all-reduce) tells every rank
 * that all writers are done.  n_dst <= 16.                                                                          */
FT_API int ft_render_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                                   float length, uint32_t flags, int32_t strip_width, int32_t rank, int32_t world,
                                   float *const *d_dst, int32_t n_dst, void *stream, ft_render_stats *stats);
/* device memory that can be shared between the per-GPU processes of one node (cudaMalloc + CUDA IPC)               */
#define FT_IPC_HANDLE_BYTES 64
FT_API int ft_device_alloc(int device, int64_t bytes, void **d_ptr);
FT_API int ft_device_free(int device, void *d_ptr);
FT_API int ft_ipc_export(int device, void *d_ptr, uint8_t handle[FT_IPC_HANDLE_BYTES]);
FT_API int ft_ipc_import(int device, const uint8_t handle[FT_IPC_HANDLE_BYTES], void **d_ptr);
FT_API int ft_ipc_close(int device, void *d_ptr);

/* number of strips and of floats in the compact per-rank buffer above       */
FT_API int ft_strip_layout(int32_t width, int32_t height, int32_t strip_width, int32_t world,
                           int32_t *strips_per_rank, int64_t *floats_per_rank);

/* all-gathered [world][strips_per_rank][strip_width][height][3] -> FColor[X,Y]
 * (device pointers, asynchronous on stream)                                  */
FT_API int ft_deinterleave_device(const float *d_gathered, int32_t width, int32_t height,
                                  int32_t strip_width, int32_t world, int32_t channels,
                                  float *d_out, void *stream);

/* the same for an all-gathered chunk [world][n_local_strips][strip_width][height][channels] that holds local
 * strips first_local_strip .. of every rank                                                                 */
FT_API int ft_deinterleave_strips_device(const float *d_gathered, int32_t width, int32_t height, int32_t strip_width,
                                         int32_t world, int32_t channels, int32_t first_local_strip,
                                         int32_t n_local_strips, float *d_out, void *stream);

/* Batched SdfScene.trace scene ray (SdfScene.fs:7-28) for arbitrary rays.
 * HOST pointers; every output is optional.  out_leaf: as above.  out_steps:
 * Distance evaluations of the primary march (SdfForm.fs:93-104).
 * out_shadow: bit l set when light l contributed (lit and unshadowed).      */
FT_API int ft_trace_rays(ft_scene *scene, int64_t n, const ft_ray *rays, uint32_t flags,
                         float *out_rgb, int32_t *out_leaf, int32_t *out_steps,
                         uint32_t *out_shadow, ft_render_stats *stats);

/* scene.Object.Form.Distance position for n positions (xyz interleaved, HOST)
 * -- the closure every march step calls (Types.fs:42).                      */
FT_API int ft_distance(ft_scene *scene, int64_t n, const float *positions, uint32_t flags,
                       float *out_distance);

/* ---- EXTENSION: spectral / refractive ray-test (BASELINE config 3) ------------------------------------------
 * No reference implementation exists: builder-defined spec in DESIGN.md 4.6, anchored in
 * the reference's dead code Light.fresnel (Light.fs:30-59) and Light.Colors (Light.fs:12-26).  n_wavelengths
 * frequencies per pixel (midpoints of equal strata of 400..789 THz), glass interactions chosen by
 * Philox4x32-10(seed; pixel, wavelength, bounce), at most max_bounces of them per path.  Wavefront pipeline on the
 * device; out_rgb is a HOST FColor[X,Y].  stats->launches counts the kernels, stats->shadow_rays every secondary
 * ray (shadow rays + continuation path segments), stats->reserved the pixels whose primary ray hit glass.       */
#define FT_SPECTRAL_MAX_LIGHTS 8
FT_API int ft_render_spectral(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                              float length, int32_t n_wavelengths, int32_t max_bounces, uint64_t seed,
                              float *out_rgb, ft_render_stats *stats);

/* the same for the strips `rank` of `world` owns, every finished pixel stored at its final position in each of the n_dst
 * destination framebuffers (DEVICE pointers, see ft_render_gather_device).  Synchronous on `stream` (the pipeline sizes its
 * path queues from a count read back after wave 0).  strip_width == 0 deals 4 x 8 pixel tiles round-robin instead of
 * column strips (tile t, counted down the columns of tiles, belongs to rank t % world): the finest balance, which these
 * workloads need -- a glass or cloud pixel costs tens of times a solid one.                                             */
FT_API int ft_render_spectral_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height,
                                            float epsilon, float length, int32_t n_wavelengths, int32_t max_bounces,
                                            uint64_t seed, int32_t strip_width, int32_t rank, int32_t world,
                                            float *const *d_dst, int32_t n_dst, void *stream, ft_render_stats *stats);

/* ---- EXTENSION: volumetric cloud material (BASELINE config 4) ----------------------------------------------------
 * No reference implementation exists: builder-defined spec in DESIGN.md 4.7.  The density field
 * is the reference's dead-code lattice noise Noise.gradient (Noise.fs:78-110, float64, used as written including the
 * `z+fz` of Noise.fs:80); its two tables are DATA passed by the caller: `perm` (256 entries, the role of
 * Noise.permutationTable, Noise.fs:7-23) and `gradients` (256 x 3 doubles, Noise.gradientTable, Noise.fs:72-76).
 * spp stochastic samples per pixel (sub-pixel jitter, fixed-step tracking with Bernoulli collisions, single
 * scattering with tracked visibility), every draw from Philox4x32-10(seed; pixel, sample, step, stream).
 * out_rgb is a HOST FColor[X,Y] holding the mean of the samples (summed in a fixed order, FT_CLOUD_SAMPLE_GROUPS below:
 * deterministic).
 * stats->shadow_rays counts every secondary march (shadow rays, continuation segments, light tracks),
 * stats->march_steps the tracking steps, stats->reserved the samples that scattered inside a cloud.            */
typedef struct ft_cloud_params {
    float density;   /* collision rate per unit length at full density                                   */
    float frequency; /* the noise is sampled at position * frequency                                    */
    float step;      /* fixed tracking step                                                             */
    float fade;      /* depth below the surface at which the density reaches its full value             */
    float bias;      /* local density = density * clamp01(bias + gain * noise) * clamp01(-Distance/fade) */
    float gain;
    const int32_t *perm;     /* 256 values in 0..255                                                    */
    const double *gradients; /* 256 x {x, y, z}                                                         */
} ft_cloud_params;
#define FT_CLOUD_MAX_SEGMENTS 8     /* camera-ray segments per sample (a segment ends where the ray leaves a cloud) */
#define FT_CLOUD_MAX_TRACK_STEPS 4096
/* a pixel's samples are summed in FT_CLOUD_SAMPLE_GROUPS consecutive groups (group g = samples [g spp / G, (g+1) spp / G)),
 * each in sample order, and the group sums are added in group order: (((P0 + P1) + P2) + P3) / spp -- so that the groups of
 * one tile can run on different warps (a tile full of cloud pixels is otherwise tens of milliseconds of serial work)       */
#define FT_CLOUD_SAMPLE_GROUPS 4
FT_API int ft_render_cloud(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height, float epsilon,
                           float length, const ft_cloud_params *cloud, int32_t samples_per_pixel, uint64_t seed,
                           float *out_rgb, ft_render_stats *stats);

/* the same for the strips `rank` of `world` owns into n_dst destination framebuffers (see ft_render_gather_device);
 * returns after the kernel has finished.                                                                            */
FT_API int ft_render_cloud_gather_device(ft_scene *scene, const ft_camera *camera, int32_t width, int32_t height,
                                         float epsilon, float length, const ft_cloud_params *cloud,
                                         int32_t samples_per_pixel, uint64_t seed, int32_t strip_width, int32_t rank,
                                         int32_t world, float *const *d_dst, int32_t n_dst, void *stream,
                                         ft_render_stats *stats);

/* The same two steps for a frame cut over several GPUs: every rank reduces and tonemaps only
 * the columns of the strips it owns (strip s -> rank s % world) of the full-size FColor[X,Y] it holds; the ranks'
 * maxima meet in an all-reduce(MAX) (Array2D.maxWith, Image.fs:40-43), and since every byte of the bitmap is written
 * by exactly one rank, an all-reduce(MAX) over the zero-initialised bitmaps assembles it
 * (fraytracer_b200/distributed.py DistributedRenderer.to_bitmap).  world = 1 equals the whole-frame calls.         */
FT_API int ft_image_max_strips_device(const float *d_rgb, int32_t width, int32_t height, int32_t strip_width,
                                      int32_t rank, int32_t world, float *d_max, void *stream);
FT_API int ft_tonemap_strips_device(const float *d_rgb, int32_t width, int32_t height, int32_t strip_width,
                                    int32_t rank, int32_t world, float gamma, const float *d_max, float max_value,
                                    uint64_t seed, uint8_t *d_out_bgr, int32_t stride, void *stream);

/* ---- post-process: Image.toColors / toBitmap / saveBitmap (Image.fs:37-90) -- */
/* max over the n_floats values of an FColor image (Array2D.Parallel.maxWith FColor.getMaxColor, Image.fs:40-42);
 * d_max is one DEVICE float.  With several GPUs, all-reduce(max) it before ft_tonemap_device.               */
FT_API int ft_image_max_device(const float *d_rgb, int64_t n_floats, float *d_max, void *stream);
/* Image.toColors gamma rng image + Image.toBitmap's pixel order, on the device.  d_rgb is FColor[X,Y]
 * (x-major).  Output: top-down 24-bit B,G,R scanlines of `stride` bytes; scanline r, column c holds the colour
 * of image[X-1-c, r] (the reference's Array.rev undoes its own y flip and mirrors x, Image.fs:66-72).
 * The maximum is read from d_max when non-NULL, else max_value is used; Max(0.01, .) is applied as Image.fs:43.
 * The reference's dither comes from a System.Random shared by a parallel map (a race: not reproducible);
 * here it is Philox4x32-10(seed, pixel index), deterministic.                                               */
FT_API int ft_tonemap_device(const float *d_rgb, int32_t width, int32_t height, float gamma, const float *d_max,
                             float max_value, uint64_t seed, uint8_t *d_out_bgr, int32_t stride, void *stream);
/* host-pointer convenience: upload FColor[X,Y], max, tone map, download the scanlines (device ordinal given) */
FT_API int ft_to_bitmap(int device, const float *rgb, int32_t width, int32_t height, float gamma, uint64_t seed,
                        uint8_t *out_bgr, int32_t stride);
/* Image.saveBitmap path (Image.fs:87-90): 24-bit BMP from top-down B,G,R scanlines (host pointers, no GPU)  */
FT_API int ft_save_bitmap(const char *path, const uint8_t *bgr, int32_t width, int32_t height, int32_t stride);

/* ---- host memory helpers ------------------------------------------------- */
FT_API int ft_host_register(void *ptr, int64_t bytes);
FT_API int ft_host_unregister(void *ptr);
FT_API int ft_host_alloc(void **ptr, int64_t bytes);
FT_API int ft_host_free(void *ptr);
/* device-side address of page-locked, mapped host memory (ft_host_alloc / ft_host_register): a valid destination of
 * ft_render_gather_device -- the kernel then stores finished pixels straight into host memory over PCIe.  With one
 * process per GPU, a framebuffer in shared host memory registered by every process lets each GPU deliver its own
 * strips over its own PCIe link.                                                                                 */
FT_API int ft_host_get_device_pointer(void *host_ptr, void **d_ptr);

/* ---- measurement --------------------------------------------------------- */
/* dependent-free FFMA stream over all SMs: the FP32 roofline denominator
 * (MEASURED_PEAKS.json carries none).  TFLOP/s, FMA counted as 2.           */
FT_API int ft_measure_fp32_peak(int device, float *tflops_fma, float *sm_mhz_estimate);

#ifdef __cplusplus
}
#endif
#endif /* FRAYTRACER_B200_H */
