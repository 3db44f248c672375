"""The host scene compiler (product) against the oracle's independent build of the same description:
spatial lookups must be identical cell by cell (centres, candidate order, LowerBounds), because the grid is
part of the definition of SdfForm.union's Distance."""
import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfMaterial, SdfObject, SdfScene, scenes
from fraytracer_b200 import _abi as A
from oracle import oracle as O

from conftest import bits

P = SdfForm.Primitive


def assert_same_grids(scene):
    tr = ft.SceneTracer(scene, device=A.FT_DEVICE_NONE)
    g1, g2 = tr.grids(), O.OracleScene(tr.description).grids()
    assert len(g1) == len(g2)
    for a, b in zip(g1, g2):
        assert a["count"] == b["count"]
        for k in ("aabb_min", "cell_size", "cell_size_inv", "centers", "lower_bound"):
            assert np.array_equal(bits(a[k]), bits(b[k])), k
        assert np.array_equal(a["begin"], b["begin"]) and np.array_equal(a["item"], b["item"])
    return tr, g1


@pytest.mark.parametrize("make", [scenes.default_scene, scenes.spheres_capsules_scene, scenes.mixed_scene])
def test_reference_configs(make):
    tr, g = assert_same_grids(make().scene)
    assert len(g) == 1


def test_nested_unions_and_shared_forms():
    b = scenes.SceneBuilder(seed=36)
    inner1 = SdfObject.union([b.randomSphere() for _ in range(7)])
    inner2 = SdfObject.union([b.randomCapsule() for _ in range(9)] + [b.randomTriangle() for _ in range(5)])
    outer = SdfObject.union([inner1, inner2, b.randomTorus(), b.randomTorus()])
    carved = SdfObject.subtract(SdfObject.intersect(outer, [P.sphere(Center=(0, 0, 0), Radius=3.0)]),
                                SdfForm.union([P.sphere(Center=(1, 1, 1), Radius=1.0), P.sphere(Center=(-1, 1, 0), Radius=0.7)]))
    scene = SdfScene(Object=carved, BackgroundColor=(0, 0, 0), Lights=scenes.default_lights())
    tr, g = assert_same_grids(scene)
    assert len(g) == 4 and tr.info()["max_depth"] == 4


def test_degenerate_grids():
    m = SdfMaterial.createSolid((1, 1, 1))
    # two identical spheres: ties in LowerBound everywhere
    s = [SdfObject.create(m, P.sphere(Center=(0, 0, 0), Radius=1.0)) for _ in range(2)]
    assert_same_grids(SdfScene(Object=SdfObject.union(s), BackgroundColor=(0, 0, 0), Lights=[]))
    # many duplicates (> 16 entries per cell exercises the introsort partition with ties)
    s = [SdfObject.create(m, P.sphere(Center=(i % 3, 0, 0), Radius=1.0)) for i in range(40)]
    assert_same_grids(SdfScene(Object=SdfObject.union(s), BackgroundColor=(0, 0, 0), Lights=[]))
    # one tiny and one huge primitive: count is clamped to >= 1 cell
    s = [SdfObject.create(m, P.sphere(Center=(0, 0, 0), Radius=100.0)), SdfObject.create(m, P.sphere(Center=(1, 0, 0), Radius=0.001))]
    tr, g = assert_same_grids(SdfScene(Object=SdfObject.union(s), BackgroundColor=(0, 0, 0), Lights=[]))
    assert g[0]["count"] == (3, 3, 3)


def test_info_counts():
    w = scenes.mixed_scene()
    tr = ft.SceneTracer(w.scene, device=A.FT_DEVICE_NONE)
    i = tr.info()
    assert i["n_primitives"] == 1024 and i["n_nodes"] == 1025 and i["n_grids"] == 1 and i["max_depth"] == 1
    assert i["n_cells"] == 12 ** 3 and i["n_entries"] == 187646
