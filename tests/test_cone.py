"""EXTENSION tests: the capped-cone primitive (BASELINE.json names "spheres + cones"; the reference has no cone,
SURVEY.md section 0, so parity is against this repository's own CPU restatement, pinned here by hand-derived distances)."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfMaterial, SdfObject, SdfScene, scenes
from fraytracer_b200 import _abi as A
from oracle import oracle as O

from conftest import bits

P = SdfForm.Primitive
GREY = SdfMaterial.createSolid((0.5, 0.5, 0.5))


def _oracle(form):
    desc = ft.SceneDescription(SdfScene(Object=SdfObject.create(GREY, form), BackgroundColor=(0.1, 0.1, 0.1), Lights=[]))
    return O.OracleScene(desc)


def test_pointed_cone_known_distances():
    orc = _oracle(P.cone(From=(0, 0, 0), To=(0, 2, 0), FromRadius=1.0, ToRadius=0.0))
    pts = np.array([[0, 3, 0],      # above the apex: 1
                    [0, -1, 0],     # below the base disc: 1
                    [3, 0, 0],      # beside the base rim (1,0,0): 2
                    [4, -4, 0],     # diagonal from the rim: 5
                    [0.5, 1, 0],    # on the slanted side: 0
                    [0, 0.5, 0]],   # inside: nearest surface is the base, -0.5
                   np.float32)
    d = orc.form_distance(0, pts)
    np.testing.assert_allclose(d, [1, 1, 2, 5, 0, -0.5], atol=2e-7)
    # a point inside near the side: distance to the line 2x + y = 2 is |2x + y - 2| / sqrt(5)
    d = orc.form_distance(0, np.array([[0.6, 0.5, 0]], np.float32))
    np.testing.assert_allclose(d, [-(2 - 1.2 - 0.5) / np.sqrt(5)], rtol=1e-6)


def test_cylinder_limit_and_boundary():
    orc = _oracle(P.cone(From=(0, 0, 0), To=(0, 0, 2), FromRadius=0.5, ToRadius=0.5))
    pts = np.array([[2, 0, 1], [0, 0, 3], [2, 0, 3], [0, 0, 1]], np.float32)
    np.testing.assert_allclose(orc.form_distance(0, pts), [1.5, 1.0, np.hypot(1.5, 1.0), -0.5], rtol=1e-6)
    b = orc.form_boundary(0)
    assert b[:3].tolist() == [0, 0, 1]
    np.testing.assert_allclose(b[3], np.hypot(1.0, 0.5), rtol=2e-6)


def test_distance_never_undercuts_the_bounding_sphere():
    """the union's pruning (SdfForm.fs:29-31) assumes Distance(p) >= |p - C| - R: holds for the cone's boundary"""
    rng = np.random.default_rng(3)
    for _ in range(20):
        a = rng.normal(size=3)
        d = rng.normal(size=3)
        form = P.cone(From=a, To=a + d / np.linalg.norm(d) * rng.uniform(0.2, 2.0), FromRadius=rng.uniform(0, 0.8),
                      ToRadius=rng.uniform(0, 0.8))
        orc = _oracle(form)
        pts = (rng.normal(size=(4000, 3)) * 2 + a).astype(np.float32)
        dist = orc.form_distance(0, pts)
        b = orc.form_boundary(0)
        bd = np.linalg.norm(pts.astype(np.float64) - b[:3], axis=1) - b[3]
        assert (dist >= bd - 1e-6).all()
        # 1-Lipschitz along random segments (an exact distance field)
        q = pts + (rng.normal(size=pts.shape) * 0.05).astype(np.float32)
        assert (np.abs(orc.form_distance(0, q) - dist) <= np.linalg.norm(q - pts, axis=1) + 1e-5).all()


def test_degenerate_cones_are_rejected():
    for form in (P.cone(From=(0, 0, 0), To=(0, 0, 0), FromRadius=1.0, ToRadius=0.5),
                 P.cone(From=(0, 0, 0), To=(0, 1, 0), FromRadius=-1.0, ToRadius=0.5)):
        scene = SdfScene(Object=SdfObject.create(GREY, form), BackgroundColor=(0.1, 0.1, 0.1), Lights=[])
        with pytest.raises(ft.FrayTracerError) as ei:
            ft.SceneTracer(scene, device=A.FT_DEVICE_NONE)
        assert ei.value.code == A.FT_ERR_INVALID_ARGUMENT


def test_compiler_grid_matches_oracle_on_the_cone_scene():
    w = scenes.spheres_cones_scene(n_spheres=60, n_cones=60)
    tr = ft.SceneTracer(w.scene, device=A.FT_DEVICE_NONE)
    g, og = tr.grids()[0], O.OracleScene(tr.description).grids()[0]
    assert g["count"] == og["count"] and np.array_equal(g["begin"], og["begin"]) and np.array_equal(g["item"], og["item"])
    assert np.array_equal(bits(g["lower_bound"]), bits(og["lower_bound"]))
    tr.close()


@pytest.mark.gpu
def test_gpu_spheres_cones_scene_bit_exact():
    w = scenes.spheres_cones_scene()
    tracer = ft.SceneTracer(w.scene, device=0)
    X, Y = 256, 144
    rgb, leaf, stats = tracer.render(w.camera, ft.ImageSize(X, Y), w.epsilon, w.length, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(w.camera, X, Y, w.epsilon, w.length, counters=True)
    assert np.array_equal(leaf, ref["leaf"]) and np.array_equal(bits(rgb), bits(ref["rgb"]))
    assert stats["shadow_rays"] == ref["counters"]["shadow_rays"] and stats["capped_rays"] == 0
    pts = (np.random.default_rng(1).normal(size=(50000, 3)) * 3).astype(np.float32)
    assert np.array_equal(bits(tracer.distance(pts)), bits(O.OracleScene(tracer.description).distance(pts)))
    tracer.close()


@pytest.mark.gpu
def test_gpu_cones_inside_composites_bit_exact():
    cone = P.cone(From=(-1, -1, 0), To=(1.5, 1, 0.5), FromRadius=1.2, ToRadius=0.2)
    obj = SdfObject.subtract(
        SdfObject.intersect(SdfObject.union([SdfObject.create(GREY, cone),
                                             SdfObject.create(SdfMaterial.createSolid((0.9, 0.2, 0.1)),
                                                              P.cone(From=(0, 2, 0), To=(0, -2, 1), FromRadius=0.0, ToRadius=0.9))]),
                            [P.sphere(Center=(0, 0, 0), Radius=2.0)]),
        P.cone(From=(0, 0, -3), To=(0, 0, 0), FromRadius=0.8, ToRadius=0.3))
    scene = SdfScene(Object=obj, BackgroundColor=(0.1, 0.2, 0.3), Lights=scenes.default_lights())
    tracer = ft.SceneTracer(scene, device=0)
    cam = scenes.default_camera()
    rgb, leaf, _ = tracer.render(cam, ft.ImageSize(150, 110), 0.01, 30.0, want_leaf=True)
    ref = O.OracleScene(tracer.description).render(cam, 150, 110, 0.01, 30.0)
    assert (leaf >= 0).any()
    assert np.array_equal(leaf, ref["leaf"]) and np.array_equal(bits(rgb), bits(ref["rgb"]))
    tracer.close()
