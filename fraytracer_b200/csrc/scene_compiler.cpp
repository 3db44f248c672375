// scene_compiler.cpp -- see scene_compiler.h.  Host code, float32, one rounding per operation
// (compiled with -ffp-contract=off).  Reference paths are relative to /root/reference/src/FrayTracer.
#include "scene_compiler.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <limits>
#include <thread>

#include "dotnet_sort.h"
#include "ft_math.h"

namespace ftc {
using namespace ftm;

namespace {

struct Bnd {
    V3 c;
    float r;
};

// SdfBoundary.fs:62-63
inline float bnd_min_distance(const Bnd &b, V3 p) { return distance(b.c, p) - b.r; }
inline float bnd_max_distance(const Bnd &b, V3 p) { return distance(b.c, p) + b.r; }

// SdfBoundary.fs:7-22
Bnd bnd_union(const Bnd &a, const Bnd &b) {
    V3 diff = b.c - a.c;
    float dist = length(diff);
    if (dist + b.r <= a.r) return a;
    if (dist + a.r <= b.r) return b;
    V3 dir = diff / dist;
    V3 a2 = a.c - dir * a.r;
    V3 b2 = b.c + dir * b.r;
    return Bnd{(a2 + b2) * 0.5f, distance(a2, b2) * 0.5f};
}

// SdfBoundary.fs:29-49 -- radius expression kept exactly as the reference writes it
Bnd bnd_intersection(const Bnd &a, const Bnd &b) {
    V3 diff = b.c - a.c;
    float dist = length(diff);
    if (dist + b.r <= a.r) return b;
    if (dist + a.r <= b.r) return a;
    V3 dir = diff / dist;
    V3 a2 = a.c + dir * a.r;
    V3 b2 = b.c - dir * b.r;
    float d2 = dist * dist;
    float aR2 = a.r * a.r;
    float bR2 = b.r * b.r;
    return Bnd{(a2 + b2) * 0.5f, sqrtf(4.0f * d2 * aR2 - (d2 - bR2 + aR2)) / (2.0f * dist)};
}

using ftsort::KeyItem;
struct DotNetSort {  // see dotnet_sort.h
    static void sort(KeyItem *a, int n) { ftsort::sort(a, n); }
};

struct FormNode {
    int kind = 0;  // FT_FORM_*
    Bnd b{};
    std::vector<int> kids;  // compiled node indices
    float strength = 0.f;
    // primitive captures
    V3 v[13] = {};
    float s[3] = {0, 0, 0};
    int grid = -1;
    int depth = 0;
    uint32_t word = 0;
};

struct ObjNode {
    int kind = 0;
    int form = -1;  // compiled node index of the object's Form
    int material = -1;
    int leaf_id = -1;
    int base = -1;
    std::vector<int> kids;  // compiled object indices
};

inline float as_float(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline uint32_t as_uint(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float4 f4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float4 f4(V3 v, float w) { return float4{v.x, v.y, v.z, w}; }

class Compiler {
   public:
    Compiler(const ft_scene_desc *d, CompiledScene *out, GridBuildFn fn, void *ctx, long long min_pairs)
        : d_(d), out_(out), grid_fn_(fn), grid_ctx_(ctx), gpu_min_pairs_(min_pairs) {}

    int run(std::string *err) {
        int st = validate();
        if (st == FT_OK) st = build();
        if (st == FT_OK) st = emit();
        if (st != FT_OK && err) *err = err_;
        return st;
    }

   private:
    const ft_scene_desc *d_;
    CompiledScene *out_;
    GridBuildFn grid_fn_;
    void *grid_ctx_;
    long long gpu_min_pairs_;
    std::string err_;
    std::vector<FormNode> nodes_;
    std::vector<ObjNode> objs_;
    std::vector<int> form_node_;  // desc form id -> compiled node (single-child composites alias the child)
    std::vector<int> obj_node_;   // desc object id -> compiled object
    std::vector<GridHost> grids_;

    int fail(int code, const std::string &msg) {
        err_ = msg;
        return code;
    }

    int validate() {
        if (!d_) return fail(FT_ERR_INVALID_ARGUMENT, "null scene description");
        if (d_->n_forms < 0 || d_->n_objects <= 0 || d_->n_children < 0 || d_->n_materials < 0 || d_->n_lights < 0)
            return fail(FT_ERR_INVALID_ARGUMENT, "negative or empty table size in scene description");
        if ((d_->n_forms && !d_->forms) || !d_->objects || (d_->n_children && !d_->children) ||
            (d_->n_materials && !d_->materials) || (d_->n_lights && !d_->lights))
            return fail(FT_ERR_INVALID_ARGUMENT, "null table pointer in scene description");
        if (d_->n_lights > FT_MAX_LIGHTS) return fail(FT_ERR_INVALID_ARGUMENT, "too many lights (FT_MAX_LIGHTS)");
        if (d_->root_object < 0 || d_->root_object >= d_->n_objects)
            return fail(FT_ERR_INVALID_ARGUMENT, "root_object out of range");
        return FT_OK;
    }

    // ---- grid: SdfBoundary.buildSpatialLookup (SdfBoundary.fs:225-274)
    int build_grid(int node_index, const std::vector<Bnd> &b) {
        const int n = (int)b.size();
        GridHost g;
        g.node = node_index;
        V3 amin = b[0].c - v3(b[0].r, b[0].r, b[0].r);  // :229 AABB.getMin
        V3 amax = b[0].c + v3(b[0].r, b[0].r, b[0].r);  // :230 AABB.getMax
        float sum = b[0].r;
        for (int i = 1; i < n; i++) {
            amin = vmin(amin, b[i].c - v3(b[i].r, b[i].r, b[i].r));
            amax = vmax(amax, b[i].c + v3(b[i].r, b[i].r, b[i].r));
            sum = sum + b[i].r;  // :232
        }
        float count_size = 1.5f * (sum / (float)n);  // :233
        V3 size = amax - amin;                       // :235
        int cx = std::max(1, ceiling_to_int(size.x / count_size));  // :237
        int cy = cx, cz = cx;  // :238-239 -- the reference divides aabbSize.X for all three axes
        if ((int64_t)cx * cy * cz > ((int64_t)1 << 27))
            return fail(FT_ERR_INVALID_ARGUMENT, "spatial lookup would need more than 2^27 cells");
        V3 cell = size / v3((float)cx, (float)cy, (float)cz);  // :240
        V3 inv = v3(1.0f, 1.0f, 1.0f) / cell;                  // :241
        g.count[0] = cx; g.count[1] = cy; g.count[2] = cz;
        g.aabb_min[0] = amin.x; g.aabb_min[1] = amin.y; g.aabb_min[2] = amin.z;
        g.cell_size[0] = cell.x; g.cell_size[1] = cell.y; g.cell_size[2] = cell.z;
        g.cell_size_inv[0] = inv.x; g.cell_size_inv[1] = inv.y; g.cell_size_inv[2] = inv.z;
        g.n_cells = (int64_t)cx * cy * cz;
        const V3 half = cell * 0.5f;
        const float half_len = length(half);

        std::vector<std::vector<KeyItem>> lists((size_t)g.n_cells);
        std::vector<V3> centers((size_t)g.n_cells);
        std::atomic<int> next_x{0};
        auto work = [&]() {
            std::vector<KeyItem> tmp((size_t)n);
            for (;;) {
                int x = next_x.fetch_add(1);
                if (x >= cx) break;
                for (int y = 0; y < cy; y++)
                    for (int z = 0; z < cz; z++) {
                        V3 center = (amin + half) + cell * v3((float)x, (float)y, (float)z);  // :246
                        float m = bnd_max_distance(b[0], center);  // :248-252 Seq.min
                        for (int i = 1; i < n; i++) {
                            float dd = bnd_max_distance(b[i], center);
                            if (dd < m) m = dd;
                        }
                        float upper = m + half_len;  // :253
                        int k = 0;
                        for (int i = 0; i < n; i++) {  // :255-264
                            float md = bnd_min_distance(b[i], center);
                            if (md < upper) tmp[(size_t)k++] = KeyItem{md, i};
                        }
                        DotNetSort::sort(tmp.data(), k);  // :267-268
                        size_t ci = ((size_t)x * cy + y) * cz + z;
                        centers[ci] = center;
                        lists[ci].assign(tmp.begin(), tmp.begin() + k);
                    }
            }
        };
        bool on_gpu = false;
        if (grid_fn_ && (long long)g.n_cells * n >= gpu_min_pairs_) {  // device build: bit-identical lists (grid_build.cu)
            std::vector<float> flat((size_t)4 * n);
            for (int i = 0; i < n; i++) {
                flat[4 * (size_t)i] = b[(size_t)i].c.x; flat[4 * (size_t)i + 1] = b[(size_t)i].c.y;
                flat[4 * (size_t)i + 2] = b[(size_t)i].c.z; flat[4 * (size_t)i + 3] = b[(size_t)i].r;
            }
            GridBuildJob job;
            job.bounds = flat.data(); job.n = n; job.cx = cx; job.cy = cy; job.cz = cz;
            job.amin[0] = amin.x; job.amin[1] = amin.y; job.amin[2] = amin.z;
            job.cell[0] = cell.x; job.cell[1] = cell.y; job.cell[2] = cell.z;
            job.half[0] = half.x; job.half[1] = half.y; job.half[2] = half.z;
            job.half_len = half_len;
            GridBuildResult res;
            std::string gerr;
            const int st = grid_fn_(grid_ctx_, job, &res, &gerr);
            if (st != FT_OK) return fail(st, gerr);
            for (int64_t c = 0; c < g.n_cells; c++) {
                centers[(size_t)c] = v3(res.centers[3 * (size_t)c], res.centers[3 * (size_t)c + 1], res.centers[3 * (size_t)c + 2]);
                const GridKeyItem *e = res.entries.data() + res.begin[(size_t)c];
                const size_t k = (size_t)(res.begin[(size_t)c + 1] - res.begin[(size_t)c]);
                lists[(size_t)c].resize(k);
                for (size_t i = 0; i < k; i++) lists[(size_t)c][i] = KeyItem{e[i].key, e[i].item};
            }
            on_gpu = true;
            out_->grids_built_on_gpu++;
        }
        unsigned hw = std::thread::hardware_concurrency();
        int nthreads = (int)std::min<unsigned>(hw ? hw : 1, 16);
        if ((int64_t)g.n_cells * n < (1 << 18)) nthreads = 1;
        if (on_gpu) {
        } else if (nthreads <= 1) {
            work();
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; t++) th.emplace_back(work);
            for (auto &t : th) t.join();
        }

        g.cell_base = (int64_t)out_->cells.size();
        g.begin.resize((size_t)g.n_cells + 1);
        int64_t off = 0;
        for (int64_t c = 0; c < g.n_cells; c++) {
            const auto &l = lists[(size_t)c];
            if (l.empty())
                return fail(FT_ERR_INVALID_ARGUMENT,
                            "spatial lookup cell without candidates (non-finite boundary?): the reference would "
                            "throw IndexOutOfRangeException in SdfForm.union");
            g.begin[(size_t)c] = off;
            off += (int64_t)l.size();
        }
        g.begin[(size_t)g.n_cells] = off;
        g.n_entries = off;
        g.lower_bound.resize((size_t)off);
        g.item.resize((size_t)off);
        if ((int64_t)out_->entries.size() + off + g.n_cells >= ((int64_t)1 << 31))
            return fail(FT_ERR_INVALID_ARGUMENT, "spatial lookup has too many candidate entries");
        for (int64_t c = 0; c < g.n_cells; c++) {
            const auto &l = lists[(size_t)c];
            uint32_t dev_begin = (uint32_t)out_->entries.size();
            out_->cells.push_back(f4(centers[(size_t)c], as_float(dev_begin)));
            int64_t o = g.begin[(size_t)c];
            for (size_t i = 0; i < l.size(); i++) {
                g.lower_bound[(size_t)o + i] = l[i].key;
                g.item[(size_t)o + i] = l[i].item;
                // child words are patched in emit(), once every node has its offset
                out_->entries.push_back(uint2{as_uint(l[i].key), (uint32_t)l[i].item});
                out_->entry_child.push_back(l[i].item);
            }
            out_->entries.push_back(uint2{as_uint(std::numeric_limits<float>::infinity()), 0u});  // sentinel
            out_->entry_child.push_back(-1);
        }
        nodes_[(size_t)node_index].grid = (int)grids_.size();
        grids_.push_back(std::move(g));
        return FT_OK;
    }

    // SdfForm.union / subtract / intersect / unionSmooth over compiled children; returns node index in *res
    int make_composite(int kind, const std::vector<int> &kids, float strength, int *res) {
        if (kind != FT_FORM_SUBTRACT) {
            if (kids.empty())  // SdfForm.fs:16,53,71
                return fail(FT_ERR_EMPTY_COMPOSITE, kind == FT_FORM_UNION_SMOOTH ? "blub" : "No SdfObjects given.");
            if (kids.size() == 1) {  // SdfForm.fs:17,54,72
                *res = kids[0];
                return FT_OK;
            }
        } else if (kids.size() != 2) {
            return fail(FT_ERR_INVALID_ARGUMENT, "SdfForm.subtract takes exactly two forms");
        }
        FormNode nd;
        nd.kind = kind;
        nd.kids = kids;
        nd.strength = strength;
        int depth = 0;
        for (int k : kids) depth = std::max(depth, nodes_[(size_t)k].depth);
        nd.depth = depth + 1;
        if (nd.depth > FT_MAX_DEPTH) return fail(FT_ERR_DEPTH, "composite forms nested deeper than FT_MAX_DEPTH");
        if (kind == FT_FORM_SUBTRACT) {
            nd.b = nodes_[(size_t)kids[0]].b;  // SdfForm.fs:48
        } else if (kind == FT_FORM_INTERSECT) {
            Bnd bb = nodes_[(size_t)kids[0]].b;  // SdfForm.fs:66 Seq.reduce
            for (size_t i = 1; i < kids.size(); i++) bb = bnd_intersection(bb, nodes_[(size_t)kids[i]].b);
            nd.b = bb;
        } else {
            Bnd bb = nodes_[(size_t)kids[0]].b;  // SdfForm.fs:36-39, 87-90 Seq.reduce
            for (size_t i = 1; i < kids.size(); i++) bb = bnd_union(bb, nodes_[(size_t)kids[i]].b);
            nd.b = bb;
        }
        nodes_.push_back(nd);
        int idx = (int)nodes_.size() - 1;
        if (kind == FT_FORM_UNION) {
            std::vector<Bnd> bb;
            bb.reserve(kids.size());
            for (int k : kids) bb.push_back(nodes_[(size_t)k].b);
            int st = build_grid(idx, bb);  // SdfForm.fs:19
            if (st != FT_OK) return st;
        }
        *res = idx;
        return FT_OK;
    }

    int build_form(int id, int depth, int *res) {
        if (id < 0 || id >= d_->n_forms) return fail(FT_ERR_INVALID_ARGUMENT, "form id out of range");
        if (form_node_[(size_t)id] >= 0) {
            *res = form_node_[(size_t)id];
            return FT_OK;
        }
        if (depth > 64) return fail(FT_ERR_DEPTH, "form nesting too deep (cycle?)");
        const ft_form_desc &f = d_->forms[id];
        const float *p = f.p;
        int idx = -1;
        switch (f.kind) {
            case FT_FORM_SPHERE: {  // SdfForm.fs:125-135
                FormNode nd;
                nd.kind = f.kind;
                nd.v[0] = v3(p[0], p[1], p[2]);
                nd.s[0] = p[3];
                nd.b = Bnd{nd.v[0], p[3]};
                nodes_.push_back(nd);
                idx = (int)nodes_.size() - 1;
                break;
            }
            case FT_FORM_CAPSULE: {  // SdfForm.fs:145-170
                FormNode nd;
                nd.kind = f.kind;
                V3 from = v3(p[0], p[1], p[2]), to = v3(p[3], p[4], p[5]);
                nd.v[0] = from;
                nd.v[1] = to - from;               // :148 dir
                nd.v[2] = inverse_length(nd.v[1]);  // :149 dirInv
                nd.s[0] = p[6];
                nd.b = Bnd{lerp(from, to, 0.5f), p[6] + distance(from, to) * 0.5f};  // :166-169
                nodes_.push_back(nd);
                idx = (int)nodes_.size() - 1;
                break;
            }
            case FT_FORM_TORUS: {  // SdfForm.fs:181-203
                FormNode nd;
                nd.kind = f.kind;
                nd.v[0] = v3(p[0], p[1], p[2]);
                nd.v[1] = normalize(v3(p[3], p[4], p[5]));  // :184
                nd.s[0] = -(dot(nd.v[0], nd.v[1]));          // :188 planeD
                nd.s[1] = p[6];
                nd.s[2] = p[7];
                nd.b = Bnd{nd.v[0], p[6] + p[7]};  // :199-202
                nodes_.push_back(nd);
                idx = (int)nodes_.size() - 1;
                break;
            }
            case FT_FORM_CONE: {  // EXTENSION: capped cone, no reference implementation
                FormNode nd;
                nd.kind = f.kind;
                V3 from = v3(p[0], p[1], p[2]), to = v3(p[3], p[4], p[5]);
                const float ra = p[6], rb = p[7];
                nd.v[0] = from;
                nd.v[1] = to - from;
                nd.s[0] = ra;
                nd.s[1] = rb;
                nd.s[2] = dot(nd.v[1], nd.v[1]);  // baba
                if (!(nd.s[2] > 0.f) || !(ra >= 0.f) || !(rb >= 0.f))
                    return fail(FT_ERR_INVALID_ARGUMENT, "cone: From and To must differ and the radii must be >= 0");
                const float h = distance(from, to) * 0.5f, rm = ra > rb ? ra : rb;
                nd.b = Bnd{lerp(from, to, 0.5f), sqrtf(h * h + rm * rm) * 1.000001f};  // rim of the larger cap, a hair of slack
                out_->has_cone = 1;
                nodes_.push_back(nd);
                idx = (int)nodes_.size() - 1;
                break;
            }
            case FT_FORM_TRIANGLE: {  // SdfForm.fs:214-263
                FormNode nd;
                nd.kind = f.kind;
                V3 v1 = v3(p[0], p[1], p[2]), v2 = v3(p[3], p[4], p[5]), v3_ = v3(p[6], p[7], p[8]);
                float R = p[9];
                V3 v21 = v2 - v1, v32 = v3_ - v2, v13 = v1 - v3_;
                V3 nor = normalize(cross(v21, v13));
                nd.v[0] = v1; nd.v[1] = v2; nd.v[2] = v3_;
                nd.v[3] = v21; nd.v[4] = v32; nd.v[5] = v13;
                nd.v[6] = inverse_length(v21); nd.v[7] = inverse_length(v32); nd.v[8] = inverse_length(v13);
                nd.v[9] = nor;
                nd.v[10] = normalize(cross(v21, nor));
                nd.v[11] = normalize(cross(v32, nor));
                nd.v[12] = normalize(cross(v13, nor));
                nd.s[0] = R;
                float area_inv = 0.5f / length2(cross(v1 - v2, v2 - v3_));                // :255
                float w1 = length2(v2 - v3_) * dot(v1 - v2, v1 - v3_) * area_inv;         // :256
                float w2 = length2(v1 - v3_) * dot(v2 - v1, v2 - v3_) * area_inv;         // :257
                float w3 = 1.0f - w1 - w2;                                                // :258
                nd.b.c = (v1 * w1 + v2 * w2) + v3_ * w3;                                  // :259
                nd.b.r = length(v21) * length(v32) * length(v13) / 2.0f / length(cross(v21, v32)) + R;  // :261-262
                nodes_.push_back(nd);
                idx = (int)nodes_.size() - 1;
                break;
            }
            case FT_FORM_UNION:
            case FT_FORM_SUBTRACT:
            case FT_FORM_INTERSECT:
            case FT_FORM_UNION_SMOOTH: {
                if (f.child_count < 0 || f.first_child < 0 || (int64_t)f.first_child + f.child_count > d_->n_children)
                    return fail(FT_ERR_INVALID_ARGUMENT, "form child range out of bounds");
                std::vector<int> kids((size_t)f.child_count);
                for (int i = 0; i < f.child_count; i++) {
                    int st = build_form(d_->children[f.first_child + i], depth + 1, &kids[(size_t)i]);
                    if (st != FT_OK) return st;
                }
                int st = make_composite(f.kind, kids, p[0], &idx);
                if (st != FT_OK) return st;
                break;
            }
            default:
                return fail(FT_ERR_UNSUPPORTED_NODE, "unsupported form kind " + std::to_string(f.kind) +
                                                         " (user-written closures cannot run on the device)");
        }
        form_node_[(size_t)id] = idx;
        *res = idx;
        return FT_OK;
    }

    int build_object(int id, int depth, int *res) {
        if (id < 0 || id >= d_->n_objects) return fail(FT_ERR_INVALID_ARGUMENT, "object id out of range");
        if (obj_node_[(size_t)id] >= 0) {
            *res = obj_node_[(size_t)id];
            return FT_OK;
        }
        if (depth > 64) return fail(FT_ERR_DEPTH, "object nesting too deep (cycle?)");
        const ft_object_desc &od = d_->objects[id];
        if (od.kind != FT_OBJECT_LEAF &&
            (od.child_count < 0 || od.first_child < 0 || (int64_t)od.first_child + od.child_count > d_->n_children))
            return fail(FT_ERR_INVALID_ARGUMENT, "object child range out of bounds");
        ObjNode o;
        o.kind = od.kind;
        o.leaf_id = id;
        switch (od.kind) {
            case FT_OBJECT_LEAF: {  // SdfObject.fs:6-10
                if (od.material < 0 || od.material >= d_->n_materials)
                    return fail(FT_ERR_INVALID_ARGUMENT, "material id out of range");
                if (d_->materials[od.material].kind != FT_MATERIAL_SOLID && d_->materials[od.material].kind != FT_MATERIAL_GLASS &&
                    d_->materials[od.material].kind != FT_MATERIAL_CLOUD)
                    return fail(FT_ERR_UNSUPPORTED_NODE, "unsupported material kind (user-written closures cannot "
                                                         "run on the device)");
                int st = build_form(od.form, 0, &o.form);
                if (st != FT_OK) return st;
                o.material = od.material;
                if (d_->materials[od.material].kind == FT_MATERIAL_GLASS) out_->has_glass = 1;
                if (d_->materials[od.material].kind == FT_MATERIAL_CLOUD) out_->has_cloud = 1;
                break;
            }
            case FT_OBJECT_UNION: {  // SdfObject.fs:12-48
                if (od.child_count == 0) return fail(FT_ERR_EMPTY_COMPOSITE, "No SdfObjects given.");  // :14
                o.kids.resize((size_t)od.child_count);
                for (int i = 0; i < od.child_count; i++) {
                    int st = build_object(d_->children[od.first_child + i], depth + 1, &o.kids[(size_t)i]);
                    if (st != FT_OK) return st;
                }
                if (od.child_count == 1) {  // :16 the object itself
                    obj_node_[(size_t)id] = o.kids[0];
                    *res = o.kids[0];
                    return FT_OK;
                }
                std::vector<int> fk;
                for (int k : o.kids) fk.push_back(objs_[(size_t)k].form);
                // :18-21 Form = SdfForm.union of the children's forms.  The material closure builds a second
                // lookup over the same boundaries (:26); buildSpatialLookup is deterministic, so the two are
                // identical and the compiled scene shares one grid.
                int st = make_composite(FT_FORM_UNION, fk, 0.f, &o.form);
                if (st != FT_OK) return st;
                break;
            }
            case FT_OBJECT_SUBTRACT: {  // SdfObject.fs:50-54
                if (od.child_count != 1) return fail(FT_ERR_INVALID_ARGUMENT, "SdfObject.subtract takes one object");
                int st = build_object(d_->children[od.first_child], depth + 1, &o.base);
                if (st != FT_OK) return st;
                int b;
                st = build_form(od.form, 0, &b);
                if (st != FT_OK) return st;
                st = make_composite(FT_FORM_SUBTRACT, {objs_[(size_t)o.base].form, b}, 0.f, &o.form);
                if (st != FT_OK) return st;
                break;
            }
            case FT_OBJECT_INTERSECT: {  // SdfObject.fs:56-64
                if (od.child_count < 1) return fail(FT_ERR_INVALID_ARGUMENT, "SdfObject.intersect needs an object");
                int st = build_object(d_->children[od.first_child], depth + 1, &o.base);
                if (st != FT_OK) return st;
                std::vector<int> fk{objs_[(size_t)o.base].form};
                for (int i = 1; i < od.child_count; i++) {
                    int k;
                    st = build_form(d_->children[od.first_child + i], 0, &k);
                    if (st != FT_OK) return st;
                    fk.push_back(k);
                }
                st = make_composite(FT_FORM_INTERSECT, fk, 0.f, &o.form);
                if (st != FT_OK) return st;
                break;
            }
            default:
                return fail(FT_ERR_UNSUPPORTED_NODE, "unsupported object kind " + std::to_string(od.kind));
        }
        objs_.push_back(o);
        obj_node_[(size_t)id] = (int)objs_.size() - 1;
        *res = obj_node_[(size_t)id];
        return FT_OK;
    }

    int build() {
        form_node_.assign((size_t)d_->n_forms, -1);
        obj_node_.assign((size_t)d_->n_objects, -1);
        for (int i = 0; i < d_->n_lights; i++)
            if (d_->lights[i].kind != FT_LIGHT_DIRECTIONAL && d_->lights[i].kind != FT_LIGHT_POINT)
                return fail(FT_ERR_UNSUPPORTED_NODE, "unsupported light kind (user-written closures cannot run on "
                                                     "the device)");
        // every constructor of the description runs, in index order: an empty composite anywhere fails at
        // construction like the reference's failwith
        int tmp;
        for (int i = 0; i < d_->n_forms; i++) {
            int st = build_form(i, 0, &tmp);
            if (st != FT_OK) return st;
        }
        for (int i = 0; i < d_->n_objects; i++) {
            int st = build_object(i, 0, &tmp);
            if (st != FT_OK) return st;
        }
        return FT_OK;
    }

    // Conservative box with Distance(p) >= dist(p, box) - slack for the form rooted at node i (device_scene.h).
    struct EscBox {
        double lo[3], hi[3], slack;
        bool ok;
    };
    EscBox escape_box(int i, int depth = 0) const {
        const FormNode &n = nodes_[(size_t)i];
        EscBox b{{0, 0, 0}, {0, 0, 0}, 0.0, true};
        if (depth > 64) { b.ok = false; return b; }
        switch (n.kind) {
            case FT_FORM_SPHERE: case FT_FORM_CAPSULE: case FT_FORM_TORUS: case FT_FORM_TRIANGLE: case FT_FORM_CONE: {
                // every primitive's surface lies inside its bounding sphere (SdfForm.fs:131-134,166-169,199-202,252-263)
                const double c[3] = {n.b.c.x, n.b.c.y, n.b.c.z}, r = n.b.r;
                b.ok = std::isfinite(r) && r >= 0.0 && std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2]);
                for (int k = 0; k < 3; k++) { b.lo[k] = c[k] - r; b.hi[k] = c[k] + r; }
                return b;
            }
            case FT_FORM_SUBTRACT:   // Max(-b, a) >= a
            case FT_FORM_INTERSECT:  // running max >= first child
                return escape_box(n.kids[0], depth + 1);
            default: {               // union: min over children; smooth union: >= min - strength*ln(n)
                for (size_t k = 0; k < n.kids.size(); k++) {
                    EscBox c = escape_box(n.kids[k], depth + 1);
                    if (!c.ok) { b.ok = false; return b; }
                    for (int a = 0; a < 3; a++) {
                        b.lo[a] = k ? std::min(b.lo[a], c.lo[a]) : c.lo[a];
                        b.hi[a] = k ? std::max(b.hi[a], c.hi[a]) : c.hi[a];
                    }
                    b.slack = std::max(b.slack, c.slack);
                }
                if (n.kind == FT_FORM_UNION_SMOOTH) {
                    if (!(n.strength > 0.f) || !std::isfinite(n.strength)) { b.ok = false; return b; }
                    b.slack += (double)n.strength * std::log((double)n.kids.size());
                }
                return b;
            }
        }
    }

    int emit() {
        CompiledScene &o = *out_;
        // 1. offsets
        std::vector<uint32_t> off(nodes_.size());
        uint32_t cur = 0;
        for (size_t i = 0; i < nodes_.size(); i++) {
            off[i] = cur;
            uint32_t kcode = 0, len = 0;
            switch (nodes_[i].kind) {
                case FT_FORM_SPHERE: kcode = FT_K_SPHERE; len = 1; break;
                case FT_FORM_CAPSULE: kcode = FT_K_CAPSULE; len = 4; break;
                case FT_FORM_TORUS: kcode = FT_K_TORUS; len = 3; break;
                case FT_FORM_TRIANGLE: kcode = FT_K_TRIANGLE; len = 14; break;
                case FT_FORM_CONE: kcode = FT_K_CONE; len = 4; break;
                case FT_FORM_UNION: kcode = FT_K_UNION; len = 4; break;
                case FT_FORM_SUBTRACT: kcode = FT_K_SUBTRACT; len = 2; break;
                case FT_FORM_INTERSECT: kcode = FT_K_INTERSECT; len = 2; break;
                case FT_FORM_UNION_SMOOTH: kcode = FT_K_SMOOTH; len = 2; break;
            }
            nodes_[i].word = FT_WORD(kcode, cur);
            cur += len;
            if (cur >= (1u << 28)) return fail(FT_ERR_INVALID_ARGUMENT, "scene too large for 28-bit node offsets");
        }
        o.nodes.assign(cur, f4(0, 0, 0, 0));
        o.n_nodes = (int)nodes_.size();
        // 2. records
        for (size_t i = 0; i < nodes_.size(); i++) {
            const FormNode &n = nodes_[i];
            float4 *r = &o.nodes[off[i]];
            r[0] = f4(n.b.c, n.b.r);
            o.max_depth = std::max(o.max_depth, n.depth);
            switch (n.kind) {
                case FT_FORM_SPHERE:
                    o.n_primitives++;
                    break;
                case FT_FORM_CAPSULE:
                    o.n_primitives++;
                    r[1] = f4(n.v[0], n.s[0]);
                    r[2] = f4(n.v[1], 0.f);
                    r[3] = f4(n.v[2], 0.f);
                    break;
                case FT_FORM_TORUS:
                    o.n_primitives++;
                    r[1] = f4(n.v[1], n.s[0]);
                    r[2] = f4(n.s[1], n.s[2], 0.f, 0.f);
                    break;
                case FT_FORM_TRIANGLE:
                    o.n_primitives++;
                    r[1] = f4(n.v[0], n.s[0]);
                    for (int k = 1; k < 13; k++) r[1 + k] = f4(n.v[k], 0.f);
                    break;
                case FT_FORM_CONE: {
                    o.n_primitives++;
                    const float rba = n.s[1] - n.s[0];
                    r[1] = f4(n.v[0], n.s[0]);
                    r[2] = f4(n.v[1], n.s[1]);
                    r[3] = f4(n.s[2], rba, rba * rba + n.s[2], 0.f);
                    break;
                }
                case FT_FORM_UNION: {
                    const GridHost &g = grids_[(size_t)n.grid];
                    r[1] = f4(g.aabb_min[0], g.aabb_min[1], g.aabb_min[2], as_float((uint32_t)g.cell_base));
                    r[2] = f4(g.cell_size_inv[0], g.cell_size_inv[1], g.cell_size_inv[2], 0.f);
                    r[3] = f4(as_float((uint32_t)(g.count[0] - 1)), as_float((uint32_t)(g.count[1] - 1)),
                              as_float((uint32_t)(g.count[2] - 1)), as_float((uint32_t)(g.count[1] * g.count[2])));
                    break;
                }
                case FT_FORM_SUBTRACT:
                    r[1] = f4(as_float(nodes_[(size_t)n.kids[0]].word), as_float(nodes_[(size_t)n.kids[1]].word), 0.f, 0.f);
                    break;
                case FT_FORM_INTERSECT:
                case FT_FORM_UNION_SMOOTH: {
                    uint32_t base = (uint32_t)o.links.size();
                    for (int k : n.kids) o.links.push_back(nodes_[(size_t)k].word);
                    if (n.kind == FT_FORM_INTERSECT)
                        r[1] = f4(as_float((uint32_t)n.kids.size()), as_float(base), 0.f, 0.f);
                    else
                        r[1] = f4(n.strength, -1.0f / n.strength /* SdfForm.fs:75 */, as_float((uint32_t)n.kids.size()),
                                  as_float(base));
                    break;
                }
            }
        }
        // 3. patch entry child words (entries hold child positions until now)
        for (const GridHost &g : grids_) {
            const FormNode &u = nodes_[(size_t)g.node];
            for (int64_t c = 0; c < g.n_cells; c++) {
                uint32_t b = as_uint(o.cells[(size_t)(g.cell_base + c)].w);
                int64_t cnt = g.begin[(size_t)c + 1] - g.begin[(size_t)c];
                for (int64_t i = 0; i < cnt; i++) {
                    uint2 &e = o.entries[(size_t)b + (size_t)i];
                    e.y = nodes_[(size_t)u.kids[(size_t)e.y]].word;
                }
            }
        }
        // 3b. chain program of the root form (device_scene.h)
        {
            std::vector<std::pair<uint32_t, uint32_t>> ops;  // collected from the root downwards
            int node = objs_[(size_t)obj_node_[(size_t)d_->root_object]].form;
            bool ok = true;
            auto is_prim = [&](int k) { return nodes_[(size_t)k].kind < FT_FORM_UNION; };
            while (ok && (nodes_[(size_t)node].kind == FT_FORM_SUBTRACT || nodes_[(size_t)node].kind == FT_FORM_INTERSECT)) {
                const FormNode &n = nodes_[(size_t)node];
                if (n.kind == FT_FORM_SUBTRACT) {
                    if (!is_prim(n.kids[1])) { ok = false; break; }
                    ops.emplace_back(FT_OP_SUBTRACT, nodes_[(size_t)n.kids[1]].word);
                } else {
                    for (size_t k = n.kids.size(); k-- > 1;) {  // reversed here, restored below
                        if (!is_prim(n.kids[k])) { ok = false; break; }
                        ops.emplace_back(FT_OP_INTERSECT, nodes_[(size_t)n.kids[k]].word);
                    }
                }
                node = n.kids[0];
            }
            if (ok && nodes_[(size_t)node].kind == FT_FORM_UNION) {
                for (int k : nodes_[(size_t)node].kids) ok = ok && is_prim(k);
                if (ok) {
                    o.chain_valid = 1;
                    o.chain_union = FT_WORD_OFF(nodes_[(size_t)node].word);
                    o.chain_n_ops = (int32_t)ops.size();
                    o.chain_ops = (uint32_t)o.links.size();
                    for (size_t k = ops.size(); k-- > 0;) {  // innermost operation first
                        o.links.push_back(ops[k].first);
                        o.links.push_back(ops[k].second);
                    }
                }
            }
        }
        if (o.links.empty()) o.links.push_back(0);
        // 4. objects
        for (const ObjNode &ob : objs_) {
            int4 rec{ob.kind, (int)nodes_[(size_t)ob.form].word, 0, 0};
            if (ob.kind == FT_OBJECT_LEAF) {
                rec.z = ob.material;
                rec.w = ob.leaf_id;
            } else if (ob.kind == FT_OBJECT_UNION) {
                rec.z = (int)o.obj_links.size();
                for (int k : ob.kids) o.obj_links.push_back(k);
            } else {
                rec.z = ob.base;
            }
            o.objects.push_back(rec);
        }
        if (o.obj_links.empty()) o.obj_links.push_back(0);
        o.root_object = obj_node_[(size_t)d_->root_object];
        if (o.chain_valid) {  // material arg-min from the hit evaluation (device_scene.h)
            int ob = o.root_object;
            int guard = 0;
            while ((objs_[(size_t)ob].kind == FT_OBJECT_SUBTRACT || objs_[(size_t)ob].kind == FT_OBJECT_INTERSECT) && guard++ < 1024)
                ob = objs_[(size_t)ob].base;
            const ObjNode &u = objs_[(size_t)ob];
            if (u.kind == FT_OBJECT_UNION && FT_WORD_OFF(nodes_[(size_t)u.form].word) == o.chain_union &&
                nodes_[(size_t)u.form].kind == FT_FORM_UNION && u.kids.size() == nodes_[(size_t)u.form].kids.size()) {
                bool leaves = true;
                for (int k : u.kids) leaves = leaves && objs_[(size_t)k].kind == FT_OBJECT_LEAF;
                if (leaves) {
                    o.chain_mat_valid = 1;
                    o.chain_mat_links = o.objects[(size_t)ob].z;
                }
            }
        }
        o.root_word = nodes_[(size_t)objs_[(size_t)o.root_object].form].word;
        // 5. materials, lights, background
        for (int i = 0; i < d_->n_materials; i++)
            o.materials.push_back(f4(d_->materials[i].rgb[0], d_->materials[i].rgb[1], d_->materials[i].rgb[2],
                                     as_float((uint32_t)d_->materials[i].kind)));
        if (o.materials.empty()) o.materials.push_back(f4(0, 0, 0, 0));
        for (int i = 0; i < d_->n_lights; i++) {
            const ft_light_desc &l = d_->lights[i];
            V3 v = v3(l.v[0], l.v[1], l.v[2]);
            if (l.kind == FT_LIGHT_DIRECTIONAL) v = normalize(-v);  // SdfLight.fs:7
            o.lights.push_back(f4(v, as_float((uint32_t)l.kind)));
            o.lights.push_back(f4(l.rgb[0], l.rgb[1], l.rgb[2], 0.f));
        }
        if (o.lights.empty()) o.lights.push_back(f4(0, 0, 0, 0));
        o.n_lights = d_->n_lights;
        o.bg[0] = d_->background[0]; o.bg[1] = d_->background[1]; o.bg[2] = d_->background[2];
        {
            const float inf = std::numeric_limits<float>::infinity();
            for (int k = 0; k < 3; k++) { o.pg_lo[k] = inf; o.pg_hi[k] = -inf; }
            int64_t best = -1;
            for (const GridHost &g : grids_)
                if (g.n_entries > best) {
                    best = g.n_entries;
                    for (int k = 0; k < 3; k++) {
                        o.pg_lo[k] = g.aabb_min[k];
                        o.pg_hi[k] = g.aabb_min[k] + g.cell_size[k] * (float)g.count[k];
                    }
                }
        }
        {
            EscBox b = escape_box(objs_[(size_t)o.root_object].form);
            o.esc_valid = b.ok ? 1 : 0;
            for (int k = 0; k < 3 && b.ok; k++) {
                // inflate: composite slack + generous room for float32 rounding of positions and distances
                const double ext = std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]));
                const double pad = b.slack + 1e-3 * (1.0 + ext + (b.hi[k] - b.lo[k]));
                o.esc_lo[k] = (float)(b.lo[k] - pad);
                o.esc_hi[k] = (float)(b.hi[k] + pad);
                if (!std::isfinite(o.esc_lo[k]) || !std::isfinite(o.esc_hi[k])) o.esc_valid = 0;
            }
        }
        o.grids = std::move(grids_);
        if (o.cells.empty()) o.cells.push_back(f4(0, 0, 0, 0));
        if (o.entries.empty()) { o.entries.push_back(uint2{as_uint(std::numeric_limits<float>::infinity()), 0}); o.entry_child.push_back(-1); }
        // 6. the 32-byte candidate records the kernels read (bounding sphere inlined) + 2 padding records
        o.fat_entries.reserve(2 * (o.entries.size() + 2));
        for (size_t e = 0; e < o.entries.size(); e++) {
            const uint2 en = o.entries[e];
            const int32_t child = o.entry_child[e];
            float4 b = f4(0, 0, 0, 0);
            if (child >= 0) b = o.nodes[FT_WORD_OFF(en.y)];
            o.fat_entries.push_back(f4(as_float(en.x), as_float(child >= 0 ? en.y : 0u), b.x, b.y));
            // .w: the radius inflated by (1 + 2^-12), rounded up -- the certain-fail filter of the sub-grid path forms
            // (m + R)(1 + 2^-12) with one fma from it (trace_device.cuh certain_fail_up)
            const float rup = std::nextafter(b.w * 1.000244140625f, std::numeric_limits<float>::infinity());
            o.fat_entries.push_back(f4(b.z, b.w, as_float((uint32_t)child), rup));
        }
        for (int k = 0; k < 2; k++) {
            o.fat_entries.push_back(f4(std::numeric_limits<float>::infinity(), 0.f, 0.f, 0.f));
            o.fat_entries.push_back(f4(0.f, 0.f, as_float((uint32_t)-1), 0.f));
        }
        return FT_OK;
    }
};

}  // namespace

int compile_scene(const ft_scene_desc *desc, CompiledScene *out, std::string *err, GridBuildFn grid_fn, void *grid_ctx,
                  long long gpu_min_pairs) {
    Compiler c(desc, out, grid_fn, grid_ctx, gpu_min_pairs);
    return c.run(err);
}

float lens_create(float field_of_view) { return (float)std::sin((double)(field_of_view * 0.5f)); }  // Camera.fs:13

void camera_look_at(const float position[3], const float look_at[3], const float up[3], float nps, ft_camera *out) {
    V3 pos = v3(position[0], position[1], position[2]);
    V3 forward = normalize(v3(look_at[0], look_at[1], look_at[2]) - pos);  // Camera.fs:34
    V3 right = normalize(cross(v3(up[0], up[1], up[2]), forward));         // Camera.fs:35
    V3 ups = cross(forward, right) * nps;                                   // Camera.fs:40
    V3 rs = right * nps;                                                    // Camera.fs:41
    out->position[0] = pos.x; out->position[1] = pos.y; out->position[2] = pos.z;
    out->forward[0] = forward.x; out->forward[1] = forward.y; out->forward[2] = forward.z;
    out->up_scaled[0] = ups.x; out->up_scaled[1] = ups.y; out->up_scaled[2] = ups.z;
    out->right_scaled[0] = rs.x; out->right_scaled[1] = rs.y; out->right_scaled[2] = rs.z;
}

}  // namespace ftc
