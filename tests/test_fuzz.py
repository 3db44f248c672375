"""Randomised scenes (random composite trees: nested unions, subtract, intersect, smooth unions over the four
primitives): the scene compiler against the oracle on the CPU, the kernels against the oracle on the GPU."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import _abi as A
from fraytracer_b200 import scenes
from oracle import oracle as O

from conftest import bits
from fuzz_scenes import random_scene

import os

# FT_FUZZ_SCENES=N: N scenes per GPU test instead of the default handful (the campaigns recorded in DESIGN.md section 5)
N_GPU = int(os.environ.get("FT_FUZZ_SCENES", "0"))


def build_both(scene, device):
    """(tracer, oracle) or (None, None) when BOTH reject the scene for the same reason: the reference's
    SdfBoundary.intersection radius formula (SdfBoundary.fs:48) yields NaN for some overlapping spheres, and a union
    over such a child has empty lookup cells -- the reference itself would throw at trace time."""
    desc = ft.SceneDescription(scene)
    try:
        orc = O.OracleScene(desc)
    except ft.FrayTracerError as e:
        assert "cell without candidates" in str(e)
        with pytest.raises(ft.FrayTracerError, match="cell without candidates"):
            ft.SceneTracer(scene, device=device)
        return None, None
    tr = ft.SceneTracer(scene, device=device)
    return tr, O.OracleScene(tr.description)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_compiler_grids_equal_oracle(seed):
    tr, orc = build_both(random_scene(seed, allow_smooth=True), A.FT_DEVICE_NONE)
    if tr is None:
        return
    g1, g2 = tr.grids(), orc.grids()
    assert len(g1) == len(g2)
    for a, b in zip(g1, g2):
        assert a["count"] == b["count"] and np.array_equal(a["item"], b["item"]) and np.array_equal(a["begin"], b["begin"])
        assert np.array_equal(bits(a["lower_bound"]), bits(b["lower_bound"])) and np.array_equal(bits(a["centers"]), bits(b["centers"]))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(N_GPU or 16))
def test_fuzz_gpu_bit_exact(seed):
    """no smooth unions (exp/log are not bit-reproducible across math libraries): everything must be bit-exact"""
    cam = scenes.default_camera()
    tracer, orc = build_both(random_scene(1000 + seed), 0)
    if tracer is None:
        return
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(20000, 3)) * 2.5).astype(np.float32)
    assert np.array_equal(bits(tracer.distance(pts)), bits(orc.distance(pts)))
    X, Y = 96, 72
    rgb, leaf, stats = tracer.render(cam, ft.ImageSize(X, Y), 0.01, 30.0, want_leaf=True)
    ref = orc.render(cam, X, Y, 0.01, 30.0, counters=True)
    assert np.array_equal(leaf, ref["leaf"]) and np.array_equal(bits(rgb), bits(ref["rgb"]))
    assert stats["shadow_rays"] == ref["counters"]["shadow_rays"] and stats["capped_rays"] == 0
    n = 3000
    origin = (rng.normal(size=(n, 3)) * 5.0).astype(np.float32)
    d = (rng.normal(size=(n, 3)) * 1.5 - origin).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([origin, d, np.full((n, 1), 20.0, np.float32), np.full((n, 1), 0.02, np.float32)], axis=1)
    got, _ = tracer.trace_rays(rays)
    want = orc.trace_rays(rays)
    for k in ("leaf", "steps", "shadow"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(bits(got["rgb"]), bits(want["rgb"]))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_fuzz_gpu_smooth_unions_within_tolerance(seed):
    tracer, orc = build_both(random_scene(2000 + seed, allow_smooth=True), 0)
    if tracer is None:
        return
    pts = (np.random.default_rng(seed).normal(size=(20000, 3)) * 2.5).astype(np.float32)
    assert np.allclose(tracer.distance(pts), orc.distance(pts), rtol=5e-5, atol=5e-6)


def _swap_materials(obj, rng, make):
    """the same object tree with about a third of its leaves' materials replaced by make()"""
    from fraytracer_b200 import SdfObject

    if obj.kind == A.FT_OBJECT_LEAF:
        return SdfObject.create(make() if rng.random() < 0.35 else obj.material, obj.form)
    if obj.kind == A.FT_OBJECT_UNION:
        return SdfObject.union([_swap_materials(c, rng, make) for c in obj.children])
    if obj.kind == A.FT_OBJECT_SUBTRACT:
        return SdfObject.subtract(_swap_materials(obj.base, rng, make), obj.form)
    return SdfObject.intersect(_swap_materials(obj.base, rng, make), list(obj.forms))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(N_GPU // 4 or 8))
def test_fuzz_gpu_extension_paths_bit_exact(seed):
    """EXTENSIONS (no reference implementation): random composite scenes with a third of the materials turned into glass /
    cloud -- the CUDA pipelines equal this repository's CPU restatements bit for bit, counters included"""
    from fraytracer_b200 import CloudParams, SdfMaterial, SdfScene

    cam = scenes.default_camera()
    base = random_scene(3000 + seed)
    rng = np.random.default_rng(seed)
    X, Y = 72, 54
    for kind in ("glass", "cloud"):
        make = (lambda: SdfMaterial.createGlass(1.45 + 0.1 * rng.random(), 0.02 + 0.03 * rng.random())) if kind == "glass" \
            else (lambda: SdfMaterial.createCloud(rng.uniform(0.5, 1.0, size=3)))
        scene = SdfScene(Object=_swap_materials(base.Object, rng, make), BackgroundColor=base.BackgroundColor, Lights=base.Lights)
        tracer, orc = build_both(scene, 0)
        if tracer is None:
            return
        if kind == "glass":
            img, st = tracer.render_spectral(cam, ft.ImageSize(X, Y), 0.01, 30.0, 8, 6, seed)
            ref = orc.render_spectral(cam, X, Y, 0.01, 30.0, 8, 6, seed)
        else:
            cp = CloudParams(density=8.0, step=0.04)
            img, st = tracer.render_cloud(cam, ft.ImageSize(X, Y), 0.01, 30.0, cp, 6, seed)
            ref, ctr = orc.render_cloud(cam, X, Y, 0.01, 30.0, cp, 6, seed)
            assert (st["shadow_rays"], st["march_steps"], st["reserved"]) == (ctr["secondary"], ctr["track_steps"], ctr["scattered"])
        assert st["capped_rays"] == 0
        assert np.array_equal(bits(img), bits(ref)), f"{kind}: {(bits(img) != bits(ref)).any(axis=2).sum()} pixels differ"
        tracer.close()
