"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on the same inputs.
Bit-exact for colours (float32), hit object index, march step counts and shadow masks."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from oracle import oracle as O

from conftest import bits

pytestmark = pytest.mark.gpu


def _check_render(w, X, Y):
    tracer = ft.SceneTracer(w.scene, device=0)
    size = ft.ImageSize(X, Y)
    rgb, leaf, stats = tracer.render(w.camera, size, w.epsilon, w.length, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length, counters=True)
    assert stats["capped_rays"] == 0
    assert np.array_equal(leaf, ref["leaf"])
    assert np.array_equal(bits(rgb), bits(ref["rgb"]))
    assert stats["primary_rays"] == X * Y
    assert stats["primary_hits"] == ref["counters"]["primary_hits"]
    assert stats["shadow_rays"] == ref["counters"]["shadow_rays"]
    tracer.close()


def test_config1_default_scene_render_bit_exact():
    _check_render(scenes.default_scene(), 192, 192)


def test_config2_spheres_capsules_render_bit_exact():
    _check_render(scenes.spheres_capsules_scene(), 256, 144)


def test_config5_mixed_render_bit_exact():
    _check_render(scenes.mixed_scene(), 256, 144)


def test_non_square_and_ragged_sizes():
    w = scenes.default_scene(n_tori=120)
    for X, Y in ((1, 1), (7, 3), (33, 65), (130, 61)):
        _check_render(w, X, Y)


def test_trace_rays_matches_oracle_per_ray():
    w = scenes.default_scene()
    tracer = ft.SceneTracer(w.scene, device=0)
    orc = O.OracleScene(tracer.description)
    rng = np.random.default_rng(7)
    n = 20000
    origin = rng.normal(size=(n, 3)).astype(np.float32) * 6.0
    target = rng.normal(size=(n, 3)).astype(np.float32) * 2.0
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([origin, d.astype(np.float32), np.full((n, 1), 25.0, np.float32),
                           rng.choice([0.01, 0.003, 0.05], size=(n, 1)).astype(np.float32)], axis=1)
    got, stats = tracer.trace_rays(rays)
    ref = orc.trace_rays(rays)
    assert np.array_equal(got["leaf"], ref["leaf"])
    assert np.array_equal(got["steps"], ref["steps"])
    assert np.array_equal(got["shadow"], ref["shadow"])
    assert np.array_equal(bits(got["rgb"]), bits(ref["rgb"]))
    assert (got["leaf"] >= 0).sum() > n // 10


def test_distance_matches_oracle():
    for w in (scenes.default_scene(), scenes.mixed_scene()):
        tracer = ft.SceneTracer(w.scene, device=0)
        orc = O.OracleScene(tracer.description)
        rng = np.random.default_rng(3)
        pts = (rng.normal(size=(50000, 3)) * 3.0).astype(np.float32)
        assert np.array_equal(bits(tracer.distance(pts)), bits(orc.distance(pts)))


def test_device_grid_equals_oracle_grid():
    w = scenes.mixed_scene()
    tracer = ft.SceneTracer(w.scene, device=0)
    g1, g2 = tracer.grids(), O.OracleScene(tracer.description).grids()
    assert len(g1) == len(g2) == 1
    for k in ("centers", "lower_bound"):
        assert np.array_equal(bits(g1[0][k]), bits(g2[0][k]))
    assert np.array_equal(g1[0]["item"], g2[0]["item"]) and np.array_equal(g1[0]["begin"], g2[0]["begin"])
