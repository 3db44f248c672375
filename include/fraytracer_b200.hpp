// fraytracer_b200.hpp -- C++17 host-side mirror of FrayTracer's F# API over the C ABI of fraytracer_b200.h.
//
// The reference is compiled code (F#, .NET) and its toolchain is absent from this repository's build image, so the host
// side above the ABI is written here in C++ (header only) with the reference's own names, argument order and failure
// messages:
//
//   reference (src/FrayTracer)                                   here (namespace fraytracer)
//   SdfForm.Primitive.sphere/capsule/torus/triangle  SdfForm.fs:125,145,181,214   SdfForm::Primitive::sphere(...) ...
//   SdfForm.union/subtract/intersect/unionSmooth     SdfForm.fs:14,42,51,69       SdfForm::union_(...) ...
//   SdfMaterial.createSolid                          SdfMaterial.fs:4-7           SdfMaterial::createSolid(color)
//   SdfObject.create/union/subtract/intersect        SdfObject.fs:6,12,50,56      SdfObject::create(material, form) ...
//   SdfLight.directional/point                       SdfLight.fs:6,23             SdfLight::directional(dir, color) ...
//   SdfScene record                                  Types.fs:74-79               SdfScene{Object, BackgroundColor, Lights}
//   SdfScene.trace scene                             SdfScene.fs:7-28             SdfScene::trace(scene, device) -> CompiledScene
//   Lens.create, Camera.lookAt                       Camera.fs:10-14,33-42        Lens::create(fov), Camera::lookAt(...)
//   Image.render eps len size camera trace           Image.fs:26-35               Image::render(eps, len, size, camera, compiled)
//   Image.toColors |> toBitmap |> saveBitmap         Image.fs:37-90               Image::toBitmap(gamma, seed, image), saveBitmap
//   System.Random + Random.fs extensions             Random.fs:9-40               DotNetRandom (scene generation only)
//
// F# closures are opaque, so the constructors return reified nodes (shared_ptr to immutable data) instead of closures;
// SdfScene::trace flattens the graph into ft_scene_desc and compiles it onto one B200.  There is no CPU fallback: without a
// device every compute call throws FrayTracerError(FT_ERR_NO_DEVICE).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fraytracer_b200.h"

namespace fraytracer {

struct FrayTracerError : std::runtime_error {
    int code;
    FrayTracerError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int status) {
    if (status != FT_OK) throw FrayTracerError(status, ft_last_error() ? ft_last_error() : "unknown error");
}

struct Vector3 {
    float X = 0, Y = 0, Z = 0;
};
inline Vector3 operator+(Vector3 a, Vector3 b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z}; }
inline Vector3 operator*(Vector3 a, float s) { return {a.X * s, a.Y * s, a.Z * s}; }
inline Vector3 operator/(Vector3 a, float s) { return {a.X / s, a.Y / s, a.Z / s}; }
using Position = Vector3;
using Normal = Vector3;

struct FColor {  // FColor.fs:7-9: a Vector3 of R, G, B
    float R = 0, G = 0, B = 0;
    static FColor ofRGB(float r, float g, float b) { return {r, g, b}; }  // FColor.fs:32-34
};
static_assert(sizeof(FColor) == 12, "FColor[X,Y] is ft_render's out_rgb: 12 bytes per element");

// ---- System.Random(seed) of .NET (Knuth subtractive generator) + Random.fs:9-40; scene generation only
class DotNetRandom {
    int32_t sa_[56] = {};
    int inext_ = 0, inextp_ = 21;

public:
    explicit DotNetRandom(int32_t seed) {
        const int32_t MBIG = 2147483647, MSEED = 161803398;
        const int32_t sub = seed == INT32_MIN ? MBIG : (seed < 0 ? -seed : seed);
        int32_t mj = MSEED - sub, mk = 1;
        sa_[55] = mj;
        int ii = 0;
        for (int i = 1; i < 55; i++) {
            ii += 21;
            if (ii >= 55) ii -= 55;
            sa_[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += MBIG;
            mj = sa_[ii];
        }
        for (int k = 0; k < 4; k++)
            for (int i = 1; i < 56; i++) {
                int n = i + 30;
                if (n >= 55) n -= 55;
                sa_[i] = (int32_t)((uint32_t)sa_[i] - (uint32_t)sa_[1 + n]);
                if (sa_[i] < 0) sa_[i] += MBIG;
            }
    }
    int32_t Next() {
        const int32_t MBIG = 2147483647;
        int a = inext_ + 1, b = inextp_ + 1;
        if (a >= 56) a = 1;
        if (b >= 56) b = 1;
        int32_t v = (int32_t)((uint32_t)sa_[a] - (uint32_t)sa_[b]);
        if (v == MBIG) v--;
        if (v < 0) v += MBIG;
        sa_[a] = v;
        inext_ = a;
        inextp_ = b;
        return v;
    }
    double NextDouble() { return Next() * (1.0 / 2147483647.0); }
    float range_01() { return (float)NextDouble(); }                           // Random.fs:9
    float range(float lo, float hi) { return lo + range_01() * (hi - lo); }   // Random.fs:11
    Vector3 pointInBall(float radius) {                                       // Random.fs:27-32
        for (;;) {
            Vector3 v;
            v.X = range(-1.f, 1.f);
            v.Y = range(-1.f, 1.f);
            v.Z = range(-1.f, 1.f);
            if ((v.X * v.X + v.Y * v.Y) + v.Z * v.Z <= 1.0f) return v * radius;
        }
    }
    Vector3 pointOnSphere(float radius) {                                     // Random.fs:34-40
        for (;;) {
            Vector3 v;
            v.X = range(-1.f, 1.f);
            v.Y = range(-1.f, 1.f);
            v.Z = range(-1.f, 1.f);
            const float l2 = (v.X * v.X + v.Y * v.Y) + v.Z * v.Z;
            if (0.01f <= l2 && l2 <= 1.0f) return (v / (float)std::sqrt((double)l2)) * radius;  // F# sqrt on float32: via double
        }
    }
};

// ---- reified nodes
struct FormNode {
    int kind = 0;
    std::array<float, 10> p{};
    std::vector<std::shared_ptr<const FormNode>> children;
};
using Form = std::shared_ptr<const FormNode>;

namespace SdfForm {
namespace Primitive {
struct Sphere { Position Center; float Radius; };                                         // SdfForm.fs:118-123
struct Capsule { Position From, To; float Radius; };                                      // SdfForm.fs:137-143
struct Torus { Position Center; Normal Normal_; float MajorRadius, MinorRadius; };        // SdfForm.fs:172-179
struct Triangle { Position V1, V2, V3; float Radius; };                                   // SdfForm.fs:205-212
struct Cone { Position From, To; float FromRadius, ToRadius; };                           // EXTENSION, not in the reference
inline Form make(int kind, std::initializer_list<float> v) {
    auto n = std::make_shared<FormNode>();
    n->kind = kind;
    int i = 0;
    for (float x : v) n->p[(size_t)i++] = x;
    return n;
}
inline Form sphere(const Sphere &d) { return make(FT_FORM_SPHERE, {d.Center.X, d.Center.Y, d.Center.Z, d.Radius}); }
inline Form capsule(const Capsule &d) { return make(FT_FORM_CAPSULE, {d.From.X, d.From.Y, d.From.Z, d.To.X, d.To.Y, d.To.Z, d.Radius}); }
inline Form torus(const Torus &d) {
    return make(FT_FORM_TORUS, {d.Center.X, d.Center.Y, d.Center.Z, d.Normal_.X, d.Normal_.Y, d.Normal_.Z, d.MajorRadius, d.MinorRadius});
}
inline Form triangle(const Triangle &d) {
    return make(FT_FORM_TRIANGLE, {d.V1.X, d.V1.Y, d.V1.Z, d.V2.X, d.V2.Y, d.V2.Z, d.V3.X, d.V3.Y, d.V3.Z, d.Radius});
}
inline Form cone(const Cone &d) {  // EXTENSION
    return make(FT_FORM_CONE, {d.From.X, d.From.Y, d.From.Z, d.To.X, d.To.Y, d.To.Z, d.FromRadius, d.ToRadius});
}
}  // namespace Primitive
inline Form composite(int kind, std::vector<Form> forms, float p0 = 0.f) {
    auto n = std::make_shared<FormNode>();
    n->kind = kind;
    n->p[0] = p0;
    n->children = std::move(forms);
    return n;
}
inline Form union_(std::vector<Form> forms) {  // SdfForm.fs:14-40
    if (forms.empty()) throw std::invalid_argument("No SdfObjects given.");  // SdfForm.fs:16
    if (forms.size() == 1) return forms[0];                                  // SdfForm.fs:17
    return composite(FT_FORM_UNION, std::move(forms));
}
inline Form subtract(Form a, Form b) { return composite(FT_FORM_SUBTRACT, {std::move(a), std::move(b)}); }  // SdfForm.fs:42-49
inline Form intersect(std::vector<Form> forms) {  // SdfForm.fs:51-67
    if (forms.empty()) throw std::invalid_argument("No SdfObjects given.");  // SdfForm.fs:53
    if (forms.size() == 1) return forms[0];
    return composite(FT_FORM_INTERSECT, std::move(forms));
}
inline Form unionSmooth(float strength, std::vector<Form> forms) {  // SdfForm.fs:69-91
    if (forms.empty()) throw std::invalid_argument("blub");  // SdfForm.fs:71
    if (forms.size() == 1) return forms[0];
    return composite(FT_FORM_UNION_SMOOTH, std::move(forms), strength);
}
}  // namespace SdfForm

struct MaterialNode {
    int kind;
    FColor rgb;
};
using Material = std::shared_ptr<const MaterialNode>;
namespace SdfMaterial {
inline Material createSolid(FColor color) { return std::make_shared<MaterialNode>(MaterialNode{FT_MATERIAL_SOLID, color}); }  // SdfMaterial.fs:4-7
inline Material createGlass(float cauchyA, float cauchyB) {  // EXTENSION
    return std::make_shared<MaterialNode>(MaterialNode{FT_MATERIAL_GLASS, FColor{cauchyA, cauchyB, 0.f}});
}
inline Material createCloud(FColor albedo) { return std::make_shared<MaterialNode>(MaterialNode{FT_MATERIAL_CLOUD, albedo}); }  // EXTENSION (4.7)
}  // namespace SdfMaterial

struct ObjectNode {
    int kind = 0;
    Form form;          // LEAF: its form; SUBTRACT: the subtracted form
    Material material;  // LEAF
    std::shared_ptr<const ObjectNode> base;                  // SUBTRACT / INTERSECT
    std::vector<std::shared_ptr<const ObjectNode>> children;  // UNION
    std::vector<Form> forms;                                 // INTERSECT: the extra forms
};
using Object = std::shared_ptr<const ObjectNode>;
namespace SdfObject {
inline Object create(Material material, Form form) {  // SdfObject.fs:6-10 (argument order: material, then form)
    auto o = std::make_shared<ObjectNode>();
    o->kind = FT_OBJECT_LEAF;
    o->material = std::move(material);
    o->form = std::move(form);
    return o;
}
inline Object union_(std::vector<Object> objects) {  // SdfObject.fs:12-48
    if (objects.empty()) throw std::invalid_argument("No SdfObjects given.");  // SdfObject.fs:14
    if (objects.size() == 1) return objects[0];
    auto o = std::make_shared<ObjectNode>();
    o->kind = FT_OBJECT_UNION;
    o->children = std::move(objects);
    return o;
}
inline Object subtract(Object object, Form form) {  // SdfObject.fs:50-54
    auto o = std::make_shared<ObjectNode>();
    o->kind = FT_OBJECT_SUBTRACT;
    o->base = std::move(object);
    o->form = std::move(form);
    return o;
}
inline Object intersect(Object object, std::vector<Form> forms) {  // SdfObject.fs:56-64
    auto o = std::make_shared<ObjectNode>();
    o->kind = FT_OBJECT_INTERSECT;
    o->base = std::move(object);
    o->forms = std::move(forms);
    return o;
}
}  // namespace SdfObject

struct LightNode {
    int kind;
    Vector3 v;
    FColor color;
};
namespace SdfLight {
inline LightNode directional(Vector3 direction, FColor color) { return {FT_LIGHT_DIRECTIONAL, direction, color}; }  // SdfLight.fs:6-21
inline LightNode point(Position position, FColor color) { return {FT_LIGHT_POINT, position, color}; }              // SdfLight.fs:23-42
}  // namespace SdfLight

struct Lens {  // Camera.fs:5-14
    float NearPlaneSize;
    static Lens create(float fieldOfView) { return {ft_lens_create(fieldOfView)}; }
};
struct Camera {  // Camera.fs:16-42
    ft_camera c{};
    static Camera lookAt(Position position, Position lookAt, Vector3 up, Lens lens) {
        Camera cam;
        const float p[3] = {position.X, position.Y, position.Z}, l[3] = {lookAt.X, lookAt.Y, lookAt.Z}, u[3] = {up.X, up.Y, up.Z};
        check(ft_camera_look_at(p, l, u, lens.NearPlaneSize, &cam.c));
        return cam;
    }
};
struct ImageSize {  // Image.fs:9-14
    int X, Y;
};

// ---- the flat ft_scene_desc of a scene (children before parents, one shared `children` index array)
class SceneDescription {
    std::vector<ft_form_desc> forms_;
    std::vector<ft_object_desc> objects_;
    std::vector<int32_t> children_;
    std::vector<ft_material_desc> materials_;
    std::vector<ft_light_desc> lights_;
    std::map<const FormNode *, int> form_ids_;
    std::map<const ObjectNode *, int> object_ids_;
    std::map<const MaterialNode *, int> material_ids_;
    ft_scene_desc desc_{};

    int add_form(const Form &f, int depth = 0) {
        if (!f) throw FrayTracerError(FT_ERR_UNSUPPORTED_NODE, "null form");
        auto it = form_ids_.find(f.get());
        if (it != form_ids_.end()) return it->second;
        if (depth > 256) throw std::invalid_argument("form nesting too deep");
        std::vector<int32_t> ids;
        for (const Form &k : f->children) ids.push_back(add_form(k, depth + 1));
        ft_form_desc d{};
        d.kind = f->kind;
        d.first_child = (int32_t)children_.size();
        d.child_count = (int32_t)ids.size();
        children_.insert(children_.end(), ids.begin(), ids.end());
        for (int i = 0; i < 10; i++) d.p[i] = f->p[(size_t)i];
        forms_.push_back(d);
        return form_ids_[f.get()] = (int)forms_.size() - 1;
    }
    int add_material(const Material &m) {
        auto it = material_ids_.find(m.get());
        if (it != material_ids_.end()) return it->second;
        ft_material_desc d{};
        d.kind = m->kind;
        d.rgb[0] = m->rgb.R;
        d.rgb[1] = m->rgb.G;
        d.rgb[2] = m->rgb.B;
        materials_.push_back(d);
        return material_ids_[m.get()] = (int)materials_.size() - 1;
    }
    int add_object(const Object &o, int depth = 0) {
        if (!o) throw FrayTracerError(FT_ERR_UNSUPPORTED_NODE, "null object");
        auto it = object_ids_.find(o.get());
        if (it != object_ids_.end()) return it->second;
        if (depth > 256) throw std::invalid_argument("object nesting too deep");
        ft_object_desc d{};
        d.kind = o->kind;
        d.form = d.material = -1;
        std::vector<int32_t> ids;
        switch (o->kind) {
            case FT_OBJECT_LEAF:
                d.form = add_form(o->form);
                d.material = add_material(o->material);
                break;
            case FT_OBJECT_UNION:
                for (const Object &c : o->children) ids.push_back(add_object(c, depth + 1));
                break;
            case FT_OBJECT_SUBTRACT:
                ids.push_back(add_object(o->base, depth + 1));
                d.form = add_form(o->form);
                break;
            default:  // FT_OBJECT_INTERSECT
                ids.push_back(add_object(o->base, depth + 1));
                for (const Form &f : o->forms) ids.push_back(add_form(f));
                break;
        }
        if (o->kind != FT_OBJECT_LEAF) {
            d.first_child = (int32_t)children_.size();
            d.child_count = (int32_t)ids.size();
            children_.insert(children_.end(), ids.begin(), ids.end());
        }
        objects_.push_back(d);
        return object_ids_[o.get()] = (int)objects_.size() - 1;
    }

public:
    SceneDescription(const Object &root, FColor background, const std::vector<LightNode> &lights) {
        const int r = add_object(root);
        for (const LightNode &l : lights) {
            ft_light_desc d{};
            d.kind = l.kind;
            d.v[0] = l.v.X; d.v[1] = l.v.Y; d.v[2] = l.v.Z;
            d.rgb[0] = l.color.R; d.rgb[1] = l.color.G; d.rgb[2] = l.color.B;
            lights_.push_back(d);
        }
        desc_.forms = forms_.data();
        desc_.objects = objects_.data();
        desc_.children = children_.data();
        desc_.materials = materials_.data();
        desc_.lights = lights_.data();
        desc_.n_forms = (int32_t)forms_.size();
        desc_.n_objects = (int32_t)objects_.size();
        desc_.n_children = (int32_t)children_.size();
        desc_.n_materials = (int32_t)materials_.size();
        desc_.n_lights = (int32_t)lights_.size();
        desc_.root_object = r;
        desc_.background[0] = background.R; desc_.background[1] = background.G; desc_.background[2] = background.B;
    }
    SceneDescription(const SceneDescription &) = delete;
    const ft_scene_desc *get() const { return &desc_; }
};

// ---- a scene compiled onto one B200: what `SdfScene.trace scene` returns here instead of a Ray -> FColor closure
class CompiledScene {
    ft_scene *h_ = nullptr;

public:
    CompiledScene(const Object &root, FColor background, const std::vector<LightNode> &lights, int device) {
        SceneDescription d(root, background, lights);
        check(ft_scene_create(d.get(), device, &h_));
    }
    CompiledScene(const CompiledScene &) = delete;
    CompiledScene(CompiledScene &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~CompiledScene() { ft_scene_destroy(h_); }
    ft_scene *handle() const { return h_; }
    ft_scene_info info() const {
        ft_scene_info i{};
        check(ft_scene_info_get(h_, &i));
        return i;
    }
    // batched SdfScene.trace scene ray (SdfScene.fs:7-28)
    std::vector<FColor> trace(const std::vector<ft_ray> &rays, std::vector<int32_t> *leaf = nullptr) const {
        std::vector<FColor> out(rays.size());
        if (leaf) leaf->resize(rays.size());
        check(ft_trace_rays(h_, (int64_t)rays.size(), rays.data(), 0, &out[0].R, leaf ? leaf->data() : nullptr, nullptr, nullptr, nullptr));
        return out;
    }
};

struct SdfScene {  // Types.fs:74-79
    Object Object_;
    FColor BackgroundColor;
    std::vector<LightNode> Lights;
    // SdfScene.trace scene (SdfScene.fs:7-28): the ray-test, here a scene compiled onto device `device`
    static CompiledScene trace(const SdfScene &scene, int device = 0) {
        return CompiledScene(scene.Object_, scene.BackgroundColor, scene.Lights, device);
    }
};

// page-locked, device-mapped FColor[X,Y]: ft_render then stores finished pixels straight into it (no copy)
class PinnedImage {
    FColor *p_ = nullptr;
    ImageSize size_{0, 0};

public:
    explicit PinnedImage(ImageSize size) : size_(size) {
        void *p = nullptr;
        check(ft_host_alloc(&p, (int64_t)size.X * size.Y * (int64_t)sizeof(FColor)));
        p_ = (FColor *)p;
    }
    PinnedImage(const PinnedImage &) = delete;
    ~PinnedImage() { ft_host_free(p_); }
    FColor *data() { return p_; }
    const FColor *data() const { return p_; }
    ImageSize size() const { return size_; }
    FColor &at(int x, int y) { return p_[(size_t)x * (size_t)size_.Y + (size_t)y]; }  // .NET [,]: x is the slow index
};

namespace Image {
// Image.render epsilon length imageSize camera (SdfScene.trace scene)  (Image.fs:26-35) into a caller-owned FColor[X,Y]
inline ft_render_stats render(float epsilon, float length, ImageSize size, const Camera &camera, const CompiledScene &trace, FColor *out,
                              int32_t *out_leaf = nullptr) {
    ft_render_stats st{};
    check(ft_render(trace.handle(), &camera.c, size.X, size.Y, epsilon, length, 0, &out[0].R, out_leaf, &st));
    return st;
}
inline std::vector<FColor> render(float epsilon, float length, ImageSize size, const Camera &camera, const CompiledScene &trace) {
    std::vector<FColor> image((size_t)size.X * (size_t)size.Y);
    render(epsilon, length, size, camera, trace, image.data());
    return image;
}
// Image.toColors gamma rng image |> Image.toBitmap (Image.fs:37-72) on the GPU: top-down B,G,R scanlines; `seed` keys the
// deterministic Philox dither that replaces the reference's shared (racy) System.Random
inline std::vector<uint8_t> toBitmap(float gamma, uint64_t seed, const FColor *image, ImageSize size, int device = 0) {
    std::vector<uint8_t> bgr((size_t)size.X * 3 * (size_t)size.Y);
    check(ft_to_bitmap(device, &image[0].R, size.X, size.Y, gamma, seed, bgr.data(), 3 * size.X));
    return bgr;
}
inline void saveBitmap(const std::string &path, const std::vector<uint8_t> &bgr, ImageSize size) {  // Image.fs:87-90
    check(ft_save_bitmap(path.c_str(), bgr.data(), size.X, size.Y, 3 * size.X));
}
}  // namespace Image

}  // namespace fraytracer
