"""SdfBoundary.buildSpatialLookup on the device.  The lists the GPU backend builds
(read back from device memory) must equal the oracle's bit for bit, ties included."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfMaterial, SdfObject, SdfScene, scenes
from oracle import oracle as O

from conftest import bits

pytestmark = pytest.mark.gpu
P = SdfForm.Primitive


def _same(scene, monkeypatch, expect_gpu=True):
    monkeypatch.setenv("FT_GRID_BUILD", "gpu")
    tr = ft.SceneTracer(scene, device=0)
    monkeypatch.delenv("FT_GRID_BUILD")
    assert (tr.info()["grids_built_on_gpu"] > 0) == expect_gpu
    g1, g2 = tr.grids(), O.OracleScene(tr.description).grids()
    assert len(g1) == len(g2)
    for a, b in zip(g1, g2):
        assert a["count"] == b["count"]
        for k in ("centers", "lower_bound"):
            assert np.array_equal(bits(a[k]), bits(b[k])), k
        assert np.array_equal(a["begin"], b["begin"]) and np.array_equal(a["item"], b["item"])
    return tr


def test_gpu_built_grids_equal_oracle(monkeypatch):
    for make in (scenes.default_scene, scenes.spheres_capsules_scene, scenes.mixed_scene):
        _same(make().scene, monkeypatch)


def test_gpu_built_grids_with_ties_and_nesting(monkeypatch):
    m = SdfMaterial.createSolid((1, 1, 1))
    s = [SdfObject.create(m, P.sphere(Center=(i % 3, 0, 0), Radius=1.0)) for i in range(40)]  # heavy ties, > 16 per cell
    _same(SdfScene(Object=SdfObject.union(s), BackgroundColor=(0, 0, 0), Lights=[]), monkeypatch)
    b = scenes.SceneBuilder(seed=78)
    inner = SdfObject.union([b.randomTriangle() for _ in range(20)])
    outer = SdfObject.union([inner] + [b.randomCapsule() for _ in range(30)])
    _same(SdfScene(Object=outer, BackgroundColor=(0, 0, 0), Lights=scenes.default_lights()), monkeypatch)


def test_render_with_gpu_built_grid_is_bit_exact(monkeypatch):
    w = scenes.mixed_scene()
    tr = _same(w.scene, monkeypatch)
    rgb, leaf, _ = tr.render(w.camera, ft.ImageSize(160, 90), w.epsilon, w.length, want_leaf=True)
    ref = O.OracleScene(tr.description).render(w.camera, 160, 90, w.epsilon, w.length)
    assert np.array_equal(bits(rgb), bits(ref["rgb"])) and np.array_equal(leaf, ref["leaf"])


def test_small_grids_stay_on_the_host_by_default():
    tr = ft.SceneTracer(scenes.default_scene(n_tori=100).scene, device=0)
    assert tr.info()["grids_built_on_gpu"] == 0
