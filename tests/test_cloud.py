"""EXTENSION tests (BASELINE config 4): volumetric cloud material, stochastic samples per pixel.

The reference has NO implementation of this (no media, no sampling).  CPU tests pin this
repository's own CPU restatement (`fto_render_cloud`, `fto_noise_gradient`) with properties of the spec and of the
reference's dead-code Noise.gradient; GPU tests compare the CUDA kernel with that restatement bit for bit."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import _abi as A
from fraytracer_b200 import scenes
from fraytracer_b200.api import (CloudParams, FColor, ImageSize, SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene)
from oracle import oracle as O

from conftest import bits

CLOUD = SdfMaterial.createCloud(FColor.ofRGB(0.9, 0.85, 0.8))
WHITE = [SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0)),
         SdfLight.point((-0.5, 0.0, -2.0), FColor.ofRGB(10.0, 10.0, 10.0))]


def _camera():
    return scenes.default_camera()


def _puff_scene(lights=WHITE, point_inside=False):
    objs = [SdfObject.create(CLOUD, SdfForm.Primitive.sphere(Center=(0.0, 0.0, 0.0), Radius=2.0)),
            SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.8, 0.3, 0.2)),
                             SdfForm.Primitive.sphere(Center=(0.5, 0.5, 5.0), Radius=2.5)),
            SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.2, 0.7, 0.9)),
                             SdfForm.Primitive.capsule(From=(-3.0, -2.6, -1.0), To=(3.0, -3.0, 1.0), Radius=0.4))]
    lights = list(lights)
    if point_inside:
        lights.append(SdfLight.point((0.2, 0.1, 0.0), FColor.ofRGB(1.0, 2.0, 3.0)))  # a light inside the medium
    return SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.3, 0.4, 0.6), Lights=lights)


def _nested_scene():
    """not a chain program: a cloud carved by subtract next to a nested union"""
    a = SdfObject.subtract(SdfObject.create(CLOUD, SdfForm.Primitive.sphere(Center=(0.0, 0.0, 0.0), Radius=2.2)),
                           SdfForm.Primitive.sphere(Center=(0.0, 0.0, -2.2), Radius=1.2))
    b = SdfObject.union([
        SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.9, 0.2, 0.2)),
                         SdfForm.Primitive.torus(Center=(0.0, 0.0, 0.0), Normal=(0.0, 1.0, 0.3), MajorRadius=3.0, MinorRadius=0.4)),
        SdfObject.create(CLOUD, SdfForm.Primitive.triangle(V1=(-3.0, -3.0, 3.0), V2=(3.0, -3.0, 3.0), V3=(0.0, 3.0, 3.5), Radius=0.6)),
    ])
    return SdfScene(Object=SdfObject.union([a, b]), BackgroundColor=FColor.ofRGB(0.2, 0.3, 0.5), Lights=WHITE)


# ----------------------------------------------------------------------------- CPU
def test_noise_gradient_known_answers():
    """Noise.gradient (Noise.fs:97-110) as written: at a lattice point every weight is 0 and fx0 = fy0 = fz0 = 0, so
    the value is gLattice = x*0 + y*0 + z + 0 = the z component of that lattice point's gradient (Noise.fs:80)."""
    perm, grad = CloudParams.default_tables()
    assert sorted(perm.tolist()) == list(range(256))
    np.testing.assert_allclose(np.linalg.norm(grad, axis=1), 1.0, rtol=1e-12)  # r = sqrt(1 - z^2): unit vectors

    def index(ix, iy, iz):
        return perm[(ix + perm[(iy + perm[iz & 255]) & 255]) & 255]

    for (ix, iy, iz) in ((0, 0, 0), (3, -2, 7), (-5, 250, 1000), (255, 256, -257)):
        got = O.noise_gradient(perm, grad, float(ix), float(iy), float(iz))
        assert got == grad[index(ix, iy, iz)][2]
    # halfway along x between two lattice points: smoothStep(0.5) = 0.5, lerp of the two gLattice values
    a = grad[index(2, 1, 4)]
    b = grad[index(3, 1, 4)]
    expect = (a[0] * 0.5 + a[2]) + 0.5 * (((b[0] * -0.5) + b[2]) - (a[0] * 0.5 + a[2]))
    assert abs(O.noise_gradient(perm, grad, 2.5, 1.0, 4.0) - expect) < 1e-15
    # the reference's generator: Noise.fs:31 System.Random(666); the first value-table draw is 2*NextDouble-1
    from fraytracer_b200.dotnet_random import DotNetRandom
    r = DotNetRandom(666)
    first = [r.NextDouble() for _ in range(258)]
    z0 = first[256] * 2.0 - 1.0
    assert grad[0][2] == z0


def test_zero_density_cloud_is_invisible_under_ambient_light():
    """no lights: only the ambient term (SdfScene.fs:12) is left, and no shadow ray exists that the cloud's surface could
    block (shadow rays are the reference's: any object's surface stops them, a cloud's too)"""
    scene = _puff_scene(lights=[])
    orc = O.OracleScene(ft.SceneDescription(scene))
    cp0 = CloudParams(density=0.0)
    img, ctr = orc.render_cloud(_camera(), 64, 48, 0.01, 30.0, cp0, spp=4, seed=3)
    assert ctr["scattered"] == 0 and ctr["track_steps"] > 0
    # the same solids without the cloud object, same jitter: equal up to the different march positions behind the medium
    solids = SdfScene(Object=SdfObject.union(scene.Object.children[1:]), BackgroundColor=scene.BackgroundColor, Lights=scene.Lights)
    ref, ctr2 = O.OracleScene(ft.SceneDescription(solids)).render_cloud(_camera(), 64, 48, 0.01, 30.0, cp0, spp=4, seed=3)
    assert ctr2["track_steps"] == 0
    close = np.abs(img - ref).max(axis=2) < 0.03
    assert close.mean() > 0.97


def test_cloud_render_is_deterministic_and_seeded():
    scene = _puff_scene()
    orc = O.OracleScene(ft.SceneDescription(scene))
    cp = CloudParams()
    a, ca = orc.render_cloud(_camera(), 48, 36, 0.01, 30.0, cp, spp=8, seed=1, threads=1)
    b, cb = orc.render_cloud(_camera(), 48, 36, 0.01, 30.0, cp, spp=8, seed=1, threads=5)
    c, _ = orc.render_cloud(_camera(), 48, 36, 0.01, 30.0, cp, spp=8, seed=2)
    assert np.array_equal(bits(a), bits(b)) and ca == cb
    assert not np.array_equal(bits(a), bits(c))
    assert ca["scattered"] > 0 and ca["secondary"] > ca["scattered"]
    # denser medium scatters more samples
    _, cd = orc.render_cloud(_camera(), 48, 36, 0.01, 30.0, CloudParams(density=30.0), spp=8, seed=1)
    assert cd["scattered"] > ca["scattered"]


def test_sample_mean_converges():
    """the 64-sample mean is closer to the 512-sample mean than the 4-sample mean is (stochastic estimator)"""
    scene = _puff_scene()
    orc = O.OracleScene(ft.SceneDescription(scene))
    cp = CloudParams()
    ref, _ = orc.render_cloud(_camera(), 32, 24, 0.01, 30.0, cp, spp=512, seed=9)
    e4 = np.abs(orc.render_cloud(_camera(), 32, 24, 0.01, 30.0, cp, spp=4, seed=1)[0] - ref).mean()
    e64 = np.abs(orc.render_cloud(_camera(), 32, 24, 0.01, 30.0, cp, spp=64, seed=1)[0] - ref).mean()
    assert e64 < 0.5 * e4


def test_cloud_scenes_compile_and_other_entry_points_reject_them():
    tr = ft.SceneTracer(_puff_scene(), device=A.FT_DEVICE_NONE)
    assert tr.info()["n_primitives"] == 3
    tr.close()


# ----------------------------------------------------------------------------- GPU
def _check(scene, X, Y, spp, seed, cp=None):
    cp = cp or CloudParams()
    tracer = ft.SceneTracer(scene, device=0)
    img, stats = tracer.render_cloud(_camera(), ImageSize(X, Y), 0.01, 30.0, cp, spp, seed)
    ref, ctr = O.OracleScene(tracer.description).render_cloud(_camera(), X, Y, 0.01, 30.0, cp, spp, seed)
    tracer.close()
    assert stats["capped_rays"] == 0
    assert np.array_equal(bits(img), bits(ref)), f"{(bits(img) != bits(ref)).any(axis=2).sum()} pixels differ"
    assert stats["primary_rays"] == X * Y * spp
    assert stats["shadow_rays"] == ctr["secondary"]
    assert stats["march_steps"] == ctr["track_steps"]
    assert stats["reserved"] == ctr["scattered"]
    return stats


@pytest.mark.gpu
def test_gpu_puff_bit_exact():
    for spp, seed in ((1, 1), (5, 19), (64, 7)):
        st = _check(_puff_scene(), 96, 72, spp, seed)
        assert st["reserved"] > 0


@pytest.mark.gpu
def test_gpu_point_light_inside_the_medium_and_dense_cloud():
    _check(_puff_scene(point_inside=True), 80, 60, 8, 3, CloudParams(density=25.0, step=0.03))
    _check(_puff_scene(lights=[]), 40, 30, 4, 3)


@pytest.mark.gpu
def test_gpu_config4_scene_bit_exact():
    w = scenes.cloud_scene(width=192, height=108)
    _check(w.scene, 192, 108, 16, 19)


@pytest.mark.gpu
def test_gpu_nested_generic_scene_and_ragged_sizes_bit_exact():
    for X, Y in ((1, 1), (33, 65), (70, 41)):
        _check(_nested_scene(), X, Y, 6, 5)


@pytest.mark.gpu
def test_gpu_other_entry_points_reject_cloud_scenes():
    tracer = ft.SceneTracer(_puff_scene(), device=0)
    with pytest.raises(ft.FrayTracerError) as ei:
        tracer.render(_camera(), ImageSize(16, 16), 0.01, 30.0)
    assert ei.value.code == A.FT_ERR_UNSUPPORTED_NODE
    with pytest.raises(ft.FrayTracerError) as ei:
        tracer.render_spectral(_camera(), ImageSize(16, 16), 0.01, 30.0)
    assert ei.value.code == A.FT_ERR_UNSUPPORTED_NODE
    tracer.close()


@pytest.mark.gpu
def test_gpu_config4_full_size_columns_bit_exact():
    """BASELINE config 4 at its full size (1920x1080, 64 samples per pixel): every 80th column of the GPU frame equals the
    CPU restatement bit for bit; re-rendering is idempotent"""
    w = scenes.cloud_scene()
    X, Y = w.size.X, w.size.Y
    cp = CloudParams()
    tracer = ft.SceneTracer(w.scene, device=0)
    img, st = tracer.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, 64, 19)
    again, st2 = tracer.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, 64, 19)
    assert np.array_equal(bits(img), bits(again)) and st["capped_rays"] == 0
    assert (st["shadow_rays"], st["march_steps"], st["reserved"]) == (st2["shadow_rays"], st2["march_steps"], st2["reserved"])
    assert st["primary_rays"] == X * Y * 64
    stride = 80
    ref, _ = O.OracleScene(tracer.description).render_cloud(w.camera, X, Y, w.epsilon, w.length, cp, 64, 19, x_stride=stride)
    cols = np.arange(0, X, stride)
    assert np.array_equal(bits(img[cols]), bits(ref[cols]))
    tracer.close()


@pytest.mark.gpu
def test_gpu_instrumented_variant_counts_work_and_changes_nothing():
    """ft_scene_set_debug bit 3: the COUNT kernel (bench.py's roofline numerator) returns the same image and counters and
    fills eval_flops; a tracking step costs at least its density evaluation (95 operations) plus a Distance evaluation."""
    cp = CloudParams()
    tracer = ft.SceneTracer(_puff_scene(), device=0)
    img, st = tracer.render_cloud(_camera(), ImageSize(96, 72), 0.01, 30.0, cp, 8, 19)
    tracer.set_debug(8)
    img_c, st_c = tracer.render_cloud(_camera(), ImageSize(96, 72), 0.01, 30.0, cp, 8, 19)
    tracer.set_debug(0)
    tracer.close()
    assert np.array_equal(bits(img), bits(img_c))
    assert st["eval_flops"] == 0
    for k in ("primary_rays", "shadow_rays", "march_steps", "reserved"):
        assert st[k] == st_c[k], k
    assert st_c["eval_flops"] > 100 * st_c["march_steps"]
