// console.cpp -- FrayTracer.Console's main program (src/FrayTracer.Console/Program.fs) over libfraytracer_b200,
// written against the C++ mirror of the F# API (include/fraytracer_b200.hpp): same scene (System.Random(19), 1000 random
// tori intersected with a sphere, a sphere subtracted, two lights), same camera, 1000 x 1000, epsilon 0.01, length 30,
// gamma 2.2, result.bmp.  The random helpers follow Program.fs:28-65 draw for draw (record fields in source order, the
// material after the form: F# evaluates `x |> f (g ())` left to right).
//
//   console [--size N] [--tori N] [--device D] [--out result.bmp] [--dump image.f32] [--describe]
//
// --describe compiles the scene without a device (FT_DEVICE_NONE) and prints what the scene compiler produced; every
// other mode needs a B200 -- there is no CPU fallback.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "fraytracer_b200.hpp"

using namespace fraytracer;

struct Generator {  // Program.fs:14, 28-65
    DotNetRandom rng{19};
    Material randomMaterial() {
        const float r = rng.range_01(), g = rng.range_01(), b = rng.range_01();
        return SdfMaterial::createSolid(FColor::ofRGB(r, g, b));
    }
    Object randomSphere() {
        SdfForm::Primitive::Sphere d{};
        d.Center = rng.pointInBall(4.0f);
        d.Radius = rng.range(0.3f, 1.0f);
        Form f = SdfForm::Primitive::sphere(d);
        return SdfObject::create(randomMaterial(), f);
    }
    Object randomCapsule() {
        const Vector3 center = rng.pointInBall(4.0f);
        SdfForm::Primitive::Capsule d{};
        d.From = center;
        const float len = rng.range(0.5f, 2.0f);
        d.To = center + rng.pointOnSphere(len);
        d.Radius = rng.range(0.1f, 0.3f);
        Form f = SdfForm::Primitive::capsule(d);
        return SdfObject::create(randomMaterial(), f);
    }
    Object randomTorus() {
        SdfForm::Primitive::Torus d{};
        d.Center = rng.pointInBall(4.0f);
        d.Normal_ = rng.pointOnSphere(1.0f);
        d.MajorRadius = rng.range(0.1f, 0.4f);
        d.MinorRadius = rng.range(0.1f, 0.3f);
        Form f = SdfForm::Primitive::torus(d);
        return SdfObject::create(randomMaterial(), f);
    }
    Object randomTriangle() {
        const Vector3 v1 = rng.pointInBall(4.0f);
        SdfForm::Primitive::Triangle d{};
        d.V1 = v1;
        const float l2 = rng.range(0.2f, 0.6f);
        d.V2 = v1 + rng.pointOnSphere(l2);
        const float l3 = rng.range(0.2f, 0.6f);
        d.V3 = v1 + rng.pointOnSphere(l3);
        d.Radius = rng.range(0.1f, 0.3f);
        Form f = SdfForm::Primitive::triangle(d);
        return SdfObject::create(randomMaterial(), f);
    }
};

int main(int argc, char **argv) {
    int size = 1000, tori = 1000, device = 0;
    std::string out = "result.bmp", dump;
    bool describe = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--size" && i + 1 < argc) size = atoi(argv[++i]);
        else if (a == "--tori" && i + 1 < argc) tori = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--dump" && i + 1 < argc) dump = argv[++i];
        else if (a == "--describe") describe = true;
        else {
            fprintf(stderr, "usage: console [--size N] [--tori N] [--device D] [--out result.bmp] [--dump image.f32] [--describe]\n");
            return 2;
        }
    }
    try {
        const Camera camera = Camera::lookAt({0.0f, 0.0f, -10.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f}, Lens::create(60.0f));  // Program.fs:16-22
        const ImageSize imageSize{size, size};                                                                                // Program.fs:24-26
        Generator gen;
        std::vector<Object> objects;
        for (int i = 1; i <= tori; i++) objects.push_back(gen.randomTorus());  // Program.fs:72-74
        SdfScene scene;
        scene.Object_ = SdfObject::subtract(                                   // Program.fs:69-78
            SdfObject::intersect(SdfObject::union_(objects), {SdfForm::Primitive::sphere({{0.f, 0.f, 0.f}, 3.5f})}),
            SdfForm::Primitive::sphere({{-0.5f, 1.f, -2.f}, 2.5f}));
        scene.BackgroundColor = FColor::ofRGB(0.1f, 0.1f, 0.1f);
        scene.Lights = {SdfLight::directional({-0.5f, -1.f, 1.f}, FColor::ofRGB(0.5f, 0.5f, 0.5f)),
                        SdfLight::point({-0.5f, 0.f, -2.f}, FColor::ofRGB(10.f, 0.f, 0.f))};
        const float epsilon = 0.01f;  // Program.fs:85

        if (describe) {
            CompiledScene compiled = SdfScene::trace(scene, FT_DEVICE_NONE);
            const ft_scene_info info = compiled.info();
            printf("{\"n_primitives\": %d, \"n_grids\": %d, \"n_cells\": %lld, \"n_entries\": %lld, \"max_depth\": %d}\n", info.n_primitives,
                   info.n_grids, (long long)info.n_cells, (long long)info.n_entries, info.max_depth);
            return 0;
        }
        CompiledScene compiled = SdfScene::trace(scene, device);  // scene compilation: outside the reference's Stopwatch too
        PinnedImage traced(imageSize);
        printf("Rendering...\n");                                 // Program.fs:87
        const auto t0 = std::chrono::steady_clock::now();         // Program.fs:89-94
        const ft_render_stats st = Image::render(epsilon, 30.0f, imageSize, camera, compiled, traced.data());
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("Time = %.2f sec\n", sec);                         // Program.fs:96
        printf("(%.3f ms; %lld primary + %lld shadow rays = %.0f Mrays/s; kernel %.3f ms)\n", sec * 1e3, (long long)st.primary_rays,
               (long long)st.shadow_rays, (double)(st.primary_rays + st.shadow_rays) / sec / 1e6, st.kernel_ms);
        if (!dump.empty()) {
            FILE *f = fopen(dump.c_str(), "wb");
            if (!f || fwrite(traced.data(), sizeof(FColor), (size_t)size * size, f) != (size_t)size * size) {
                fprintf(stderr, "cannot write %s\n", dump.c_str());
                return 1;
            }
            fclose(f);
        }
        Image::saveBitmap(out, Image::toBitmap(2.2f, 19, traced.data(), imageSize, device), imageSize);  // Program.fs:98-100
        return 0;
    } catch (const FrayTracerError &e) {
        fprintf(stderr, "fraytracer_b200 error %d: %s\n", e.code, e.what());
        return 1;
    }
}
