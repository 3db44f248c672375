// device_scene.h -- the compiled scene as it lives in HBM (and as the host compiler emits it).
//
// Everything is a handful of flat arrays (structure-of-arrays at the array level, 16-byte records inside
// so that every access is one vectorised LDG.128 / LDG.64 through the read-only path):
//
//   nodes[]   float4   form nodes, variable length records.  A node is addressed by a 32-bit "word":
//                      bits 31..28 = kind code (FT_K_*), bits 27..0 = index of its first float4.
//                      record[0] is always the bounding sphere {cx,cy,cz,r} (Types.fs:19-24).
//       sphere    [0] {C, R}                                  (its distance IS the bound test's value)
//       capsule   [1] {From, R} [2] {dir, -} [3] {dirInv, -}                       (SdfForm.fs:148-164)
//       torus     [1] {N, planeD} [2] {Major, Minor, -, -}      (C = bound centre)  (SdfForm.fs:182-194)
//       triangle  [1] {V1,R} [2] {V2} [3] {V3} [4] {v21} [5] {v32} [6] {v13} [7] {v21'} [8] {v32'}
//                 [9] {v13'} [10] {nor} [11] {n21} [12] {n32} [13] {n13}             (SdfForm.fs:216-250)
//       union     [1] {aabbMin, bits(cellBase)} [2] {cellSizeInv, -} [3] bits{cx-1, cy-1, cz-1, cy*cz}
//                                                                 (SdfBoundary.fs:276-282, SdfForm.fs:22-34)
//       subtract  [1] bits{word a, word b, -, -}                                    (SdfForm.fs:45-47)
//       intersect [1] bits{nchild, linkBase, -, -}    links[linkBase+i] = child word (SdfForm.fs:58-64)
//       smooth    [1] {strength, -1/strength, bits(nchild), bits(linkBase)}         (SdfForm.fs:75-82)
//   cells[]   float4   {centre.xyz, bits(entryBegin)} per grid cell, x-major like the .NET [,,] array
//   entries[] float4x2 one 32-byte record per candidate: {LowerBound, bits(child word), cx, cy} {cz, r,
//                      bits(child position), -}: the child's bounding sphere is inlined so that the two
//                      pruning tests of SdfForm.fs:29-31 need no dependent load.  Each cell's list is sorted
//                      ascending by LowerBound (SdfBoundary.fs:267-268) and closed by a sentinel {+inf, ..}:
//                      the loop `if !(min > lb - dtc) break` is exactly equivalent to the reference's full
//                      scan because the list is sorted and `min` never grows (DESIGN.md "early exit").
//                      Two padding records follow the last list (the scan prefetches one record ahead).
//   objects[] int4     {kind, form word, a, b}: LEAF a=material b=leaf id | UNION a=objLinkBase b=unused
//                      | SUBTRACT/INTERSECT a=base object
//   obj_links[] int    object index of child i of an object union
//   materials[] float4 {r,g,b,-}
//   lights[]  float4x2 {bits(kind), v.xyz} {rgb, -}   v = normalize(-direction) or the position
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
#endif

enum : uint32_t {
    FT_K_SPHERE = 0,  // 0: "is a sphere" is one unsigned comparison of the word
    FT_K_CAPSULE = 2,
    FT_K_TORUS = 3,
    FT_K_TRIANGLE = 4,
    FT_K_CONE = 5,  // EXTENSION (no reference): [1] {From, ra} [2] {To - From, rb} [3] {baba, rb - ra, (rb-ra)^2 + baba, -}
    FT_K_UNION = 8,
    FT_K_SUBTRACT = 9,
    FT_K_INTERSECT = 10,
    FT_K_SMOOTH = 11,
};
#define FT_WORD(kind, off) (((uint32_t)(kind) << 28) | (uint32_t)(off))
#define FT_WORD_KIND(w) ((w) >> 28)
#define FT_WORD_OFF(w) ((w)&0x0FFFFFFFu)
#define FT_IS_PRIMITIVE(w) (((w) >> 31) == 0u)

#define FT_DEV_MAX_LIGHTS 32

struct DevScene {
    const float4 *nodes;
    const float4 *cells;
    const float4 *entries;  // 2 float4 per record
    const uint32_t *links;
    const int4 *objects;
    const int32_t *obj_links;
    const float4 *materials;
    const float4 *lights;
    uint32_t root_word;   // form word of scene.Object.Form
    int32_t root_object;
    int32_t n_lights;
    int32_t step_cap;
    float bg[3];
    // world-space box of the largest spatial lookup.  Scheduling hint only (never used in arithmetic):
    // evaluations outside it scan whole candidate lists, inside it only a few entries.
    float pg_lo[3], pg_hi[3];
    // Escape box: a box (already inflated by a safety slack) such that scene.Object.Form.Distance(p) >=
    // dist(p, box) for every p (every primitive's bounding sphere lies inside it and the composites above
    // only increase the distance).  A ray whose remaining half-line stays farther than epsilon from it can
    // never satisfy `distance < epsilon` (SdfForm.fs:98): it is a miss without marching.  esc_valid == 0
    // disables the shortcut (smooth unions with non-positive strength, non-finite boundaries, ...).
    float esc_lo[3], esc_hi[3];
    int32_t esc_valid;
    // Chain program: when scene.Object.Form is  op_k(... op_1(U) ...)  with U a spatial-lookup union whose
    // children are all primitives and every op an SdfForm.intersect / SdfForm.subtract whose other operands
    // are primitives (the shape of the reference's own scenes, Program.fs:69-78), the form is evaluated by a
    // specialised linear evaluator instead of the generic interpreter.  links[chain_ops + 2k] = opcode
    // (FT_OP_*), links[chain_ops + 2k + 1] = primitive word, in evaluation order.
    int32_t chain_valid;
    uint32_t chain_union;  // node offset of U
    int32_t chain_n_ops;
    uint32_t chain_ops;
    // records [1..3] of U (aabbMin | cellBase, cellSizeInv, counts): kernel-parameter copies, so that the chain
    // evaluator's cell lookup starts from the constant bank instead of three dependent global loads
    float4 chain_g1, chain_g2, chain_g3;
    // When the object tree above U is subtract / intersect objects over ONE SdfObject.union whose form is U (the shape of
    // the reference's scenes), the material arg-min of SdfObject.fs:27-46 at the hit point is the candidate that set the
    // running minimum of the hit's own Distance evaluation (same grid, same order, same two pruning tests, and `d < min`
    // updates exactly when Min(min, d) changes): the chain evaluator notes it and no second scan is needed.
    // obj_links[chain_mat_links + child position] is the leaf object.
    int32_t chain_mat_valid;
    int32_t chain_mat_links;
    // Nearest-candidate sub-grid of U (subgrid_build.cu, DESIGN.md "sub-grid"): a FINE grid of sub_scale sub-cells per cell and
    // axis that also covers a halo of sub_halo cells around U's box.  With c = (p - aabbMin) * cellSizeInv (the reference's own
    // cell coordinate, SdfBoundary.fs:276), the fine index along an axis is  g = floor(c * sub_scale) + sub_off  (the product is
    // exact: sub_scale is a power of two), valid for sub_lo <= c < sub_hi; the reference's cell is clamp(floor(c)) =
    // clamp((g - sub_off) >> sub_shift), so a fine cell maps to ONE reference cell, halo included (border cells).
    // sub_hdr[(gx * sub_ny + gy) * sub_nz + gz] = {(first << 8) | count, (q << 24) | nearest candidate}, q = a lower bound of
    // dist(p, reference cell centre) over the fine cell in quanta of 1 / (256 sub_dtc_inv) (certificate (1) of sub_lookup);
    // sub_idx[first ...] holds candidates
    // 1 .. count-1: ABSOLUTE indices into entries[] -- those candidates of the reference cell's own list that can be the nearest
    // one for some position inside the fine cell (all others are provably farther than the nearest by a margin), nearest first.
    // Positions inside the covered region evaluate only these (eval_chain); the result is certified equal to the reference's
    // sequential scan, otherwise the scan itself runs.  sub_valid == 0: no sub-grid (the scan always runs).
    const uint2 *sub_hdr;
    const uint2 *sub_hdr_biased;  // sub_hdr + ((sub_off * sub_ny + sub_off) * sub_nz + sub_off): indexed by floor(c * sub_scale) directly
    const uint32_t *sub_idx;
    int32_t sub_valid;
    int32_t sub_shift;      // log2(sub-cells per cell and axis): 0 .. 3
    int32_t sub_off;        // sub_halo << sub_shift
    int32_t sub_ny, sub_nz; // fine cells along y and z
    float sub_scale;        // (float)(1 << sub_shift)
    float sub_dtc_inv;      // 1 / (256 quanta): fma(x, sub_dtc_inv, 1) <= as_float(0x3f800000 | q << 15)  <=>  x <= q quanta
    float sub_lo[3], sub_hi[3];  // -halo and count + halo per axis, as floats
    // "Fast region" in the same cell coordinates: the covered region clipped to the escape box (when there is one).  Inside it
    // the evaluator needs no further coverage test and the ray machine no escape pre-check (trace_device.cuh in_fast_region);
    // empty (lo = +inf) without a sub-grid.  Scheduling only: every position outside it takes the complete tests.
    float fast_lo[3], fast_hi[3];
    // the same in fine-cell units (times sub_scale), and cellSizeInv * sub_scale: the ray machine's per-iteration box test and
    // the fine-cell index work on  f = (p - aabbMin) * chain_g2s  directly (trace_device.cuh chain_coord)
    float fast_lo_s[3], fast_hi_s[3];
    float4 chain_g2s;
    int32_t sub_debug;      // bit 0: distrust every certificate -- every evaluation runs the reference's sequential scan (tests).
                            // Applied on the host (ft_scene_set_debug detaches the sub-grid: sub_valid = 0, empty fast region),
                            // so the kernels carry no test for it
};
enum : uint32_t {
    FT_OP_INTERSECT = 1,  // if v < dist(C,p)+R then v = Max(v, D(p))      SdfForm.fs:62-63
    FT_OP_SUBTRACT = 2,   // v = Max(-D(p), v)                             SdfForm.fs:45-47
};
