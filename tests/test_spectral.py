"""EXTENSION tests (BASELINE config 3): spectral / refractive ray-test.

The reference has NO implementation of this (no wavelengths, no glass), so nothing here is
"parity with the reference": the CPU tests pin this repository's own CPU restatement (`fto_render_spectral`) with
properties that follow from the spec, the GPU tests compare the CUDA wavefront pipeline with that restatement
bit for bit."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import _abi as A
from fraytracer_b200 import scenes
from fraytracer_b200.api import (Camera, FColor, ImageSize, Lens, SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene)
from oracle import oracle as O

from conftest import bits

GLASS = SdfMaterial.createGlass(1.5046, 0.042)


def _camera():
    return scenes.default_camera()


def _glass_ball_scene(lights=(), solid_behind=False, bg=(0.3, 0.4, 0.6)):
    objs = [SdfObject.create(GLASS, SdfForm.Primitive.sphere(Center=(0.0, 0.0, 0.0), Radius=2.0))]
    if solid_behind:
        objs.append(SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.8, 0.7, 0.2)),
                                     SdfForm.Primitive.sphere(Center=(0.5, 0.5, 4.0), Radius=2.5)))
        objs.append(SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.2, 0.7, 0.9)),
                                     SdfForm.Primitive.capsule(From=(-3.0, -1.0, -1.0), To=(3.0, -2.0, 1.0), Radius=0.4)))
    return SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(*bg), Lights=list(lights))


def _nested_scene():
    """not a chain program: a glass object carved by subtract, inside a union with a nested union and a torus"""
    a = SdfObject.subtract(SdfObject.create(GLASS, SdfForm.Primitive.sphere(Center=(0.0, 0.0, 0.0), Radius=2.2)),
                           SdfForm.Primitive.sphere(Center=(0.0, 0.0, -2.2), Radius=1.2))
    b = SdfObject.union([
        SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.9, 0.2, 0.2)),
                         SdfForm.Primitive.torus(Center=(0.0, 0.0, 0.0), Normal=(0.0, 1.0, 0.3), MajorRadius=3.0, MinorRadius=0.4)),
        SdfObject.create(SdfMaterial.createSolid(FColor.ofRGB(0.2, 0.9, 0.3)),
                         SdfForm.Primitive.triangle(V1=(-3.0, -3.0, 3.0), V2=(3.0, -3.0, 3.0), V3=(0.0, 3.0, 3.5), Radius=0.3)),
    ])
    lights = [SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0)),
              SdfLight.point((-0.5, 0.0, -6.0), FColor.ofRGB(30.0, 20.0, 10.0))]
    return SdfScene(Object=SdfObject.union([a, b]), BackgroundColor=FColor.ofRGB(0.2, 0.3, 0.5), Lights=lights)


# ----------------------------------------------------------------------------- CPU: the restatement against its spec
def test_grey_solid_scene_equals_the_reference_render():
    """without glass, with grey colours everywhere, spec(grey, nu) = grey for every band: every wavelength carries the
    reference's Lambert value, so the spectral image equals SdfScene.trace's image up to the final sum's rounding"""
    b = scenes.SceneBuilder(5)
    grey = lambda v: SdfMaterial.createSolid(FColor.ofRGB(v, v, v))
    objs = [SdfObject.create(grey(0.2 + 0.05 * (i % 10)), b.randomSphere().form) for i in range(40)]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.1, 0.1, 0.1),
                     Lights=[SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(0.5, 0.5, 0.5)),
                             SdfLight.point((-0.5, 0.0, -2.0), FColor.ofRGB(10.0, 10.0, 10.0))])
    orc = O.OracleScene(ft.SceneDescription(scene))
    ref = orc.render(_camera(), 64, 48, 0.01, 30.0)["rgb"]
    for nl in (6, 16, 64):  # every band is populated, so every channel has a positive weight sum
        img = orc.render_spectral(_camera(), 64, 48, 0.01, 30.0, n_wavelengths=nl, max_bounces=4, seed=1)
        np.testing.assert_allclose(img, ref, rtol=2e-6, atol=1e-7)


def test_glass_ball_before_a_uniform_background_is_invisible():
    """every path carries weight one and ends on the background (no lights, nothing else in the scene): the glass ball
    cannot be seen, except where a path spends its whole interaction budget (total internal reflection chains)"""
    scene = _glass_ball_scene()
    orc = O.OracleScene(ft.SceneDescription(scene))
    img = orc.render_spectral(_camera(), 96, 96, 0.01, 30.0, n_wavelengths=16, max_bounces=64, seed=3)
    bg = img[0, 0].copy()  # a pixel that misses everything: the band round trip of the background colour
    assert np.abs(bg - np.array([0.3, 0.4, 0.6])).max() < 0.25
    err = np.abs(img - bg).max(axis=2)
    assert (err < 1e-6).mean() > 0.999
    # with no interaction allowed the ball is black
    img0 = orc.render_spectral(_camera(), 96, 96, 0.01, 30.0, n_wavelengths=16, max_bounces=0, seed=3)
    ball = np.abs(img0).max(axis=2) == 0.0
    assert 0.02 < ball.mean() < 0.5
    assert np.abs(img0[~ball] - bg).max() < 1e-6


def test_fresnel_reflectance_at_normal_incidence_is_four_percent():
    """Light.fresnel with the ray-oriented normal: ((n-1)/(n+1))^2 at normal incidence.  One interaction allowed, black
    background behind, a bright one would need a second scene; instead count reflected paths: with max_bounces = 1 a
    reflected path leaves towards the camera side and ends on the background (value bg), a transmitted one hits the
    far side of the ball (glass again, budget spent -> 0)."""
    scene = _glass_ball_scene(bg=(1.0, 1.0, 1.0))
    orc = O.OracleScene(ft.SceneDescription(scene))
    X = 257
    img = orc.render_spectral(_camera(), X, X, 0.01, 30.0, n_wavelengths=64, max_bounces=1, seed=11)
    # the pixels around the ball's centre (the ball is centred on the optical axis: pixel X/2)
    c = X // 2
    patch = img[c - 6:c + 7, c - 6:c + 7, 1]  # green channel: mean over its wavelengths of "reflected?"
    n_mid = 1.5046 + 0.042 / (299.792458 / 560.0) ** 2
    expect = ((n_mid - 1) / (n_mid + 1)) ** 2
    assert abs(patch.mean() - expect) < 0.012, (patch.mean(), expect)


def test_seed_changes_only_glass_pixels_and_is_deterministic():
    scene = _glass_ball_scene(lights=[SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0))], solid_behind=True)
    orc = O.OracleScene(ft.SceneDescription(scene))
    a = orc.render_spectral(_camera(), 80, 60, 0.01, 30.0, 16, 8, seed=1, threads=1)
    b = orc.render_spectral(_camera(), 80, 60, 0.01, 30.0, 16, 8, seed=1, threads=7)
    c = orc.render_spectral(_camera(), 80, 60, 0.01, 30.0, 16, 8, seed=2)
    assert np.array_equal(bits(a), bits(b))
    black = orc.render_spectral(_camera(), 80, 60, 0.01, 30.0, 16, 0, seed=1)
    glass_px = np.abs(black).max(axis=2) == 0.0
    assert glass_px.any() and not np.array_equal(bits(a), bits(c))
    assert np.array_equal(bits(a)[~glass_px], bits(c)[~glass_px])


def test_plain_render_rejects_glass_scenes():
    scene = _glass_ball_scene()
    tr = ft.SceneTracer(scene, device=A.FT_DEVICE_NONE)  # compiles (validation only)
    assert tr.info()["n_primitives"] == 1
    tr.close()


# ----------------------------------------------------------------------------- GPU: CUDA wavefront pipeline == restatement
def _check(scene, X, Y, nl, mb, seed, eps=0.01, length=30.0):
    tracer = ft.SceneTracer(scene, device=0)
    img, stats = tracer.render_spectral(_camera(), ImageSize(X, Y), eps, length, nl, mb, seed)
    ref = O.OracleScene(tracer.description).render_spectral(_camera(), X, Y, eps, length, nl, mb, seed)
    tracer.close()
    assert stats["capped_rays"] == 0
    assert np.array_equal(bits(img), bits(ref)), f"{(bits(img) != bits(ref)).any(axis=2).sum()} pixels differ"
    return stats


@pytest.mark.gpu
def test_gpu_glass_ball_bit_exact():
    lights = [SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0)),
              SdfLight.point((-0.5, 0.0, -6.0), FColor.ofRGB(30.0, 30.0, 30.0))]
    scene = _glass_ball_scene(lights=lights, solid_behind=True)
    for nl, mb in ((1, 1), (7, 3), (16, 8), (64, 2), (16, 0)):
        st = _check(scene, 120, 90, nl, mb, seed=19)
        assert st["reserved"] > 0  # glass pixels
        assert st["launches"] == (3 + mb if mb > 0 else 3)


@pytest.mark.gpu
def test_gpu_config3_scene_bit_exact():
    w = scenes.spectral_glass_scene(width=256, height=144)
    st = _check(w.scene, 256, 144, 16, 8, seed=19)
    assert st["shadow_rays"] > st["reserved"] * 16  # continuation segments were marched


@pytest.mark.gpu
def test_gpu_nested_generic_scene_and_ragged_sizes_bit_exact():
    scene = _nested_scene()
    for X, Y in ((1, 1), (33, 65), (130, 61)):
        _check(scene, X, Y, 16, 6, seed=5)


@pytest.mark.gpu
def test_gpu_no_glass_in_view_and_solid_only():
    b = scenes.SceneBuilder(5)
    objs = [b.randomSphere() for _ in range(30)]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.1, 0.2, 0.3), Lights=scenes.default_lights())
    st = _check(scene, 97, 55, 16, 8, seed=1)
    assert st["reserved"] == 0 and st["launches"] == 1


@pytest.mark.gpu
def test_gpu_plain_render_rejects_glass_and_spectral_is_deterministic():
    scene = _glass_ball_scene(solid_behind=True, lights=scenes.default_lights())
    tracer = ft.SceneTracer(scene, device=0)
    with pytest.raises(ft.FrayTracerError) as ei:
        tracer.render(_camera(), ImageSize(16, 16), 0.01, 30.0)
    assert ei.value.code == A.FT_ERR_UNSUPPORTED_NODE
    a, _ = tracer.render_spectral(_camera(), ImageSize(200, 120), 0.01, 30.0, 16, 8, 7)
    b, _ = tracer.render_spectral(_camera(), ImageSize(200, 120), 0.01, 30.0, 16, 8, 7)
    assert np.array_equal(bits(a), bits(b))
    tracer.close()


@pytest.mark.gpu
def test_gpu_config3_full_size_columns_bit_exact():
    """BASELINE config 3 at its full size (3840x2160, 16 wavelengths): every 96th column of the GPU frame equals the CPU
    restatement bit for bit; re-rendering is idempotent; glass pixels are exactly those that go black with no bounce"""
    w = scenes.spectral_glass_scene()
    X, Y = w.size.X, w.size.Y
    tracer = ft.SceneTracer(w.scene, device=0)
    img, st = tracer.render_spectral(w.camera, w.size, w.epsilon, w.length, 16, 8, 19)
    again, _ = tracer.render_spectral(w.camera, w.size, w.epsilon, w.length, 16, 8, 19)
    assert np.array_equal(bits(img), bits(again)) and st["capped_rays"] == 0
    stride = 96
    ref = O.OracleScene(tracer.description).render_spectral(w.camera, X, Y, w.epsilon, w.length, 16, 8, 19, x_stride=stride)
    cols = np.arange(0, X, stride)
    assert np.array_equal(bits(img[cols]), bits(ref[cols]))
    black, st0 = tracer.render_spectral(w.camera, w.size, w.epsilon, w.length, 16, 0, 19)
    assert int((np.abs(black).max(axis=2) == 0.0).sum()) == st["reserved"] == st0["reserved"]
    tracer.close()


@pytest.mark.gpu
def test_gpu_instrumented_variant_counts_work_and_changes_nothing():
    """ft_scene_set_debug bit 3: the COUNT kernels (bench.py's roofline numerator) return the same image bit for bit and fill
    eval_flops / march_steps; the plain kernels leave them at zero."""
    w = scenes.spectral_glass_scene(width=160, height=90)
    tracer = ft.SceneTracer(w.scene, device=0)
    img, st = tracer.render_spectral(_camera(), ImageSize(160, 90), 0.01, 30.0, 16, 8, 19)
    tracer.set_debug(8)
    img_c, st_c = tracer.render_spectral(_camera(), ImageSize(160, 90), 0.01, 30.0, 16, 8, 19)
    tracer.set_debug(0)
    tracer.close()
    assert np.array_equal(bits(img), bits(img_c))
    assert st["eval_flops"] == 0 and st["march_steps"] == 0
    rays = st_c["primary_rays"] + st_c["shadow_rays"]
    assert st_c["shadow_rays"] == st["shadow_rays"] and st_c["reserved"] == st["reserved"]
    assert st_c["march_steps"] > rays  # several Distance evaluations per ray ...
    assert 50 * st_c["march_steps"] < st_c["eval_flops"] < 5000 * st_c["march_steps"]  # ... each tens to hundreds of flops
