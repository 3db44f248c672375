"""Pins the CPU oracle (oracle/ft_oracle.c) with known answers.

The reference ships no tests or golden vectors, so these are (a) the published
behaviour of .NET's System.Random, (b) distances/boundaries derived by hand from the reference formulas
(SdfForm.fs, SdfBoundary.fs), (c) an independent pure-Python float32 restatement of the grid union, and
(d) a regression fixture of the oracle's own output (tests/golden/, made by tests/golden/make_golden.py).
"""
import json
import math
import struct
import os

import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene, scenes
from fraytracer_b200 import _abi as A
from fraytracer_b200.dotnet_random import DotNetRandom
from oracle import oracle as O

from conftest import bits

f32 = np.float32
P = SdfForm.Primitive
GREY = SdfMaterial.createSolid((0.5, 0.5, 0.5))


def scene_of(obj):
    return SdfScene(Object=obj, BackgroundColor=(0.1, 0.1, 0.1), Lights=[])


def oracle_of(form_or_obj):
    obj = form_or_obj if isinstance(form_or_obj, SdfObject) else SdfObject.create(GREY, form_or_obj)
    desc = ft.SceneDescription(scene_of(obj))
    return O.OracleScene(desc), desc


# ----------------------------------------------------------------------------- System.Random
def test_dotnet_random_known_first_values():
    # well-known outputs of `new System.Random(seed).Next()` on every .NET version with the seeded algorithm
    for seed, first in ((0, 1559595546), (1, 534011718), (42, 1434747710)):
        assert DotNetRandom(seed).Next() == first
        assert O.OracleRandom(seed).Next() == first


def test_dotnet_random_python_and_c_agree_on_long_stream():
    for seed in (19, 666, -7, 2147483647):
        a, b = DotNetRandom(seed), O.OracleRandom(seed)
        for _ in range(3000):
            assert a.Next() == b.Next()
    a, b = DotNetRandom(19), O.OracleRandom(19)
    assert [a.NextDouble() for _ in range(4)] == [b.NextDouble() for _ in range(4)] == [
        0.6523242395614852, 0.39227909380210524, 0.04879450846872968, 0.6135768725599986]


def test_program_fs_scene_regression_anchor():
    # torus #0 of Program.fs with System.Random(19) (scratch-derived there, independent of this code)
    w = scenes.default_scene()
    t = w.leaves[0]
    want = (1.2185941, -0.8617673, -3.609644, 0.24944666, 0.6243194, -0.74027133, 0.22573131, 0.2948363)
    assert np.allclose(t.form.params, want, rtol=0, atol=5e-7)
    assert np.allclose(t.material.rgb, (0.32816705, 0.27997532, 0.2897863), atol=5e-8)
    assert len(w.leaves) == 1000


# ----------------------------------------------------------------------------- .NET sort
def test_dotnet_sort_sorts_and_is_stable_for_small_lists():
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 3, 5, 16, 17, 64, 1000):
        keys = rng.integers(0, max(2, n // 3 + 1), size=n).astype(np.float32)  # many ties
        k, v = O.dotnet_sort(keys, np.arange(n))
        assert np.all(np.diff(k) >= 0)
        assert sorted(v.tolist()) == list(range(n)) and np.array_equal(keys[v], k)
        if n <= 16 and n != 3:  # insertion sort below the introsort threshold is stable
            assert np.array_equal(v, np.argsort(keys, kind="stable"))


def test_dotnet_sort_three_elements_network_and_nan_front():
    # size 3 uses the (0,1)(0,2)(1,2) swap network: with keys (1,1,0) the tie order flips
    k, v = O.dotnet_sort([1.0, 1.0, 0.0], [0, 1, 2])
    assert v.tolist() == [2, 1, 0]
    k, v = O.dotnet_sort([3.0, np.nan, 1.0, np.nan, 2.0], [0, 1, 2, 3, 4])
    assert np.isnan(k[:2]).all() and k[2:].tolist() == [1.0, 2.0, 3.0]


# ----------------------------------------------------------------------------- primitives (hand-derived)
def test_sphere_distance_and_boundary():
    orc, _ = oracle_of(P.sphere(Center=(1, 2, 3), Radius=0.5))
    pts = np.array([[1, 2, 3], [1, 2, 5], [4, 6, 3], [1, 2.5, 3]], np.float32)
    assert orc.form_distance(0, pts).tolist() == [-0.5, 1.5, 4.5, 0.0]
    assert orc.form_boundary(0).tolist() == [1, 2, 3, 0.5]


def test_capsule_three_branches_and_boundary():
    orc, _ = oracle_of(P.capsule(From=(0, 0, 0), To=(2, 0, 0), Radius=0.25))
    pts = np.array([[-3, 4, 0],   # t <= 0: |p - From| - R = 5 - .25
                    [5, 0, 4],    # t >= 1: |p - To| - R = 5 - .25
                    [1, 3, 0],    # middle: distance to the axis = 3 - .25
                    [0.5, 0, 0]], np.float32)
    assert orc.form_distance(0, pts).tolist() == [4.75, 4.75, 2.75, -0.25]
    assert orc.form_boundary(0).tolist() == [1, 0, 0, 1.25]  # lerp(From,To,.5), R + |To-From|/2


def test_torus_distance_and_boundary():
    # normal given un-normalised (SdfForm.fs:184 normalises it)
    orc, _ = oracle_of(P.torus(Center=(0, 0, 0), Normal=(0, 0, 5), MajorRadius=2.0, MinorRadius=0.5))
    pts = np.array([[2, 0, 0],     # on the ring circle: -minor
                    [4, 0, 0],     # in plane, outside: 4-2-.5
                    [0, 0, 0],     # centre: major - minor
                    [2, 0, 3],     # above the ring: 3 - .5
                    [0, 0, 1.5]], np.float32)
    want = [-0.5, 1.5, 1.5, 2.5, float(f32(math.sqrt(1.5 * 1.5 + 4.0)) - f32(0.5))]
    assert orc.form_distance(0, pts).tolist() == want
    assert orc.form_boundary(0).tolist() == [0, 0, 0, 2.5]


def test_triangle_face_edge_vertex_and_boundary():
    orc, _ = oracle_of(P.triangle(V1=(0, 0, 0), V2=(4, 0, 0), V3=(0, 4, 0), Radius=0.5))
    pts = np.array([[1, 1, 2],     # above the face: |z| - R
                    [1, 1, -3],    # below the face
                    [2, -3, 0],    # outside edge V1V2: 3 - R
                    [-3, -4, 0],   # nearest vertex V1: 5 - R
                    [7, 0, 4]], np.float32)  # nearest vertex V2: 5 - R
    assert orc.form_distance(0, pts).tolist() == [1.5, 2.5, 2.5, 4.5, 4.5]
    b = orc.form_boundary(0)
    # circumcentre of the right triangle = midpoint of the hypotenuse, circumradius = half of it, + R
    assert np.allclose(b, [2, 2, 0, math.sqrt(32) / 2 + 0.5], atol=1e-6)


# ----------------------------------------------------------------------------- composites
def test_subtract_and_intersect_semantics():
    a = P.sphere(Center=(0, 0, 0), Radius=2.0)
    b = P.sphere(Center=(1, 0, 0), Radius=1.0)
    orc, desc = oracle_of(SdfForm.subtract(a, b))
    pts = np.array([[0, 0, 0], [1, 0, 0], [-1.5, 0, 0], [5, 0, 0]], np.float32)
    fid = desc.desc.n_forms - 1
    da = np.array([-2, -1, -0.5, 3], np.float32)
    db = np.array([0, -1, 1.5, 3], np.float32)
    assert orc.form_distance(fid, pts).tolist() == np.maximum(-db, da).tolist()  # Max(-b, a)  SdfForm.fs:45-47
    orc, desc = oracle_of(SdfForm.intersect([a, b]))
    fid = desc.desc.n_forms - 1
    # SdfForm.fs:58-64: child i>=1 only evaluated when max < dist(C_i,p)+R_i; otherwise the running max stays
    want = [float(max(x, y)) if x < (y + 2.0) else float(x) for x, y in zip(da, db)]
    assert orc.form_distance(fid, pts).tolist() == want


def test_union_smooth_formula():
    forms = [P.sphere(Center=(0, 0, 0), Radius=1.0), P.sphere(Center=(3, 0, 0), Radius=1.0),
             P.torus(Center=(0, 2, 0), Normal=(0, 1, 0), MajorRadius=1.0, MinorRadius=0.25)]
    k = 0.3
    orc, desc = oracle_of(SdfForm.unionSmooth(k, forms))
    rng = np.random.default_rng(0)
    pts = (rng.normal(size=(200, 3)) * 2).astype(np.float32)
    fid = desc.desc.n_forms - 1
    d = np.stack([orc.form_distance(i, pts) for i in range(3)]).astype(np.float64)
    want = -np.log(np.exp(-d / k).sum(0)) * k  # SdfForm.fs:75-82
    assert np.allclose(orc.form_distance(fid, pts), want, rtol=2e-5, atol=2e-6)


def test_single_child_composites_alias_the_child_and_empty_ones_fail():
    s = P.sphere(Center=(0, 0, 0), Radius=1.0)
    assert SdfForm.union([s]) is s and SdfForm.intersect([s]) is s and SdfForm.unionSmooth(0.1, [s]) is s
    for ctor in (SdfForm.union, SdfForm.intersect, SdfObject.union):
        with pytest.raises(ValueError, match="No SdfObjects given."):
            ctor([])
    with pytest.raises(ValueError, match="blub"):
        SdfForm.unionSmooth(0.1, [])
    # the same through the flat description: the oracle raises like the reference's failwith
    desc = ft.SceneDescription(scene_of(SdfObject.create(GREY, s)))
    desc._forms[0].kind = A.FT_FORM_UNION
    desc._forms[0].child_count = 0
    with pytest.raises(ft.FrayTracerError, match="No SdfObjects given."):
        O.OracleScene(desc)


# ----------------------------------------------------------------------------- independent restatement of the grid union
def _py_union_distance(prims, bounds, p):
    """SdfBoundary.buildSpatialLookup + SdfForm.union (SdfBoundary.fs:225-282, SdfForm.fs:22-34) in plain
    Python float32, brute force per query point: rebuilds the cell of p only."""
    def dist(a, b):
        d = (a - b).astype(f32)
        return f32(np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))))
    C = np.array([b[:3] for b in bounds], f32)
    R = np.array([b[3] for b in bounds], f32)
    amin = (C - R[:, None]).min(0)
    amax = (C + R[:, None]).max(0)
    s = f32(0)
    for r in R:
        s = f32(s + r)
    count_size = f32(f32(1.5) * f32(s / f32(len(R))))
    size = (amax - amin).astype(f32)
    n = max(1, int(math.ceil(f32(size[0] / count_size))))
    cell = (size / f32(n)).astype(f32)
    inv = (f32(1) / cell).astype(f32)
    c = ((p - amin).astype(f32) * inv).astype(f32)
    idx = np.clip(np.floor(c).astype(int), 0, n - 1)
    center = ((amin + (cell * f32(0.5)).astype(f32)).astype(f32) + (cell * idx.astype(f32)).astype(f32)).astype(f32)
    half = (cell * f32(0.5)).astype(f32)
    half_len = dist(half, np.zeros(3, f32))
    upper = f32(min(f32(dist(C[i], center) + R[i]) for i in range(len(R))) + half_len)
    items = [(f32(dist(C[i], center) - R[i]), i) for i in range(len(R))]
    items = sorted([it for it in items if it[0] < upper], key=lambda t: t[0])  # ties: see test below
    dtc = dist(p, center)
    m = prims[items[0][1]](p)
    for lb, i in items[1:]:
        if m > f32(lb - dtc) and m > f32(dist(C[i], p) - R[i]):
            m = min(m, prims[i](p))
    return m


def test_grid_union_against_independent_python_restatement():
    rng = np.random.default_rng(5)
    centers = (rng.normal(size=(40, 3)) * 2).astype(np.float32)
    radii = rng.uniform(0.2, 0.9, size=40).astype(np.float32)
    forms = [P.sphere(Center=c, Radius=r) for c, r in zip(centers, radii)]
    orc, desc = oracle_of(SdfForm.union(forms))
    fid = desc.desc.n_forms - 1

    def sphere_fn(c, r):
        def f(p):
            d = (c - p).astype(f32)
            return f32(f32(np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2])))) - r)
        return f
    prims = [sphere_fn(c, r) for c, r in zip(centers, radii)]
    bounds = [np.array([*c, r], f32) for c, r in zip(centers, radii)]
    pts = (rng.normal(size=(300, 3)) * 3).astype(np.float32)
    got = orc.form_distance(fid, pts)
    want = np.array([_py_union_distance(prims, bounds, p) for p in pts], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_default_scene_grid_statistics():
    # SURVEY.md section 8(a) A7 (scratch estimates of the survey, made independently of this code)
    w = scenes.default_scene()
    orc = O.OracleScene(ft.SceneDescription(w.scene))
    g = orc.grids()[0]
    lens = np.diff(g["begin"])
    assert g["count"] == (14, 14, 14) and len(g["item"]) == 152538
    assert lens.min() == 11 and lens.max() == 146 and abs(lens.mean() - 55.6) < 0.05 and np.median(lens) == 52
    for c in range(0, len(lens), 97):  # every list sorted ascending by LowerBound
        lb = g["lower_bound"][g["begin"][c]:g["begin"][c + 1]]
        assert np.all(np.diff(lb) >= 0)


# ----------------------------------------------------------------------------- camera, ray generation, shading
def test_lens_and_camera_quirks():
    nps = O.lens_create(60.0)  # sin(30 rad): the reference passes degrees to sin (Camera.fs:13, Program.fs:21)
    assert nps == float(f32(math.sin(30.0))) and abs(nps + 0.9880316) < 1e-7
    cam = O.camera_look_at((0, 0, -10), (0, 0, 0), (0, 1, 0), nps)
    assert list(cam.forward) == [0, 0, 1]
    assert list(cam.right_scaled) == [nps, 0, 0] and list(cam.up_scaled) == [0, nps, 0]
    # FSharp.Core 5.0.1 evaluates `sin` of a float32 in double and rounds once; a later FSharp.Core calling MathF.Sin would give
    # the same NearPlaneSize at the reference's own field of view: the correctly rounded float32 sine of 30 is 0xbf7cefa4 either way
    # (the double-precision sine lies far from the midpoints between neighbouring float32 values, so the float32 result does not
    # depend on how the last bits of a correctly rounded or nearly correctly rounded sine fall)
    assert struct.unpack("<I", struct.pack("<f", nps))[0] == 0xBF7CEFA4
    lo, hi = struct.unpack("<f", struct.pack("<I", 0xBF7CEFA3))[0], struct.unpack("<f", struct.pack("<I", 0xBF7CEFA5))[0]
    exact = math.sin(30.0)
    assert abs(nps - exact) < 0.25 * abs(lo - nps) and abs(nps - exact) < 0.25 * abs(hi - nps)


def test_trace_single_sphere_by_hand():
    # one sphere, one directional light shining along +z (from behind the camera)
    obj = SdfObject.create(SdfMaterial.createSolid((0.5, 0.25, 1.0)), P.sphere(Center=(0, 0, 0), Radius=1.0))
    scene = SdfScene(Object=obj, BackgroundColor=(0.1, 0.2, 0.3), Lights=[SdfLight.directional((0, 0, 1), (2, 2, 2))])
    orc = O.OracleScene(ft.SceneDescription(scene))
    rays = np.array([[0.3, 0, -5, 0, 0, 1, 30, 0.01],   # off-axis: hits, lit, unshadowed
                     [0, 3, -5, 0, 0, 1, 30, 0.01],     # misses
                     [0, 0, -5, 0, 0, 1, 2.0, 0.01],    # Length too short: miss
                     [0, 0, -5, 0, 0, 1, 30, 0.01]],    # head-on: lands exactly on the surface (d = 0)
                    np.float32)
    r = orc.trace_rays(rays)
    assert r["leaf"].tolist() == [0, -1, -1, 0] and r["shadow"].tolist() == [1, 0, 0, 0]
    assert r["steps"].tolist()[3] == 2  # d=4 (move), then d=0 < eps
    bg = np.array([0.1, 0.2, 0.3], f32)
    col = np.array([0.5, 0.25, 1.0], f32)
    pi_inv = f32(1) / f32(math.pi)
    assert bits(r["rgb"][1]).tolist() == bits(bg).tolist() == bits(r["rgb"][2]).tolist()
    # lit: colour = Color * ((bg + light*cos) * piInv), normal ~ radial, cos ~ sqrt(1 - 0.3^2)
    want = col * ((bg + f32(2.0 * math.sqrt(1 - 0.09))) * pi_inv)
    assert np.allclose(r["rgb"][0], want, rtol=2e-2)
    # head-on: the hit ray is backed off by eps along the ray (SdfObject.fs:73) to z = -1.01f, where the scene
    # distance 1.01f - 1 = 0.00999999 is already < eps: the shadow ray "hits" at once (self-shadow, SdfLight.fs:10-20).
    # Reference behaviour, kept; the colour is then exactly Color * (bg * piInv)  (SdfScene.fs:12,28)
    assert bits(r["rgb"][3]).tolist() == bits(col * (bg * pi_inv)).tolist()


def test_render_layout_is_x_major_and_matches_trace_rays():
    w = scenes.default_scene(n_tori=60, size=24)
    orc = O.OracleScene(ft.SceneDescription(w.scene))
    img = orc.render(w.camera, 24, 16, w.epsilon, w.length)
    # build the same rays by hand: Image.fs:17-23 + Camera.fs:44-54, no pixel-centre offset, max(X,Y) normalisation
    cam = w.camera
    rays = []
    for x in range(24):
        for y in range(16):
            u, v = f32(f32(x) / f32(24)), f32(f32(y) / f32(24))
            d = (np.array(cam.Forward, f32) + (u - f32(0.5)) * np.array(cam.RightScaled, f32)).astype(f32)
            d = (d + (v - f32(0.5)) * np.array(cam.UpScaled, f32)).astype(f32)
            l = f32(np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))))
            rays.append([*cam.Position, *(d / l).astype(f32), w.length, w.epsilon])
    r = orc.trace_rays(np.array(rays, f32))
    assert np.array_equal(bits(img["rgb"].reshape(-1, 3)), bits(r["rgb"]))
    assert np.array_equal(img["leaf"].reshape(-1), r["leaf"])


def test_oracle_golden_regression_fixture():
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_config1_64.json")
    g = json.load(open(path))
    w = scenes.default_scene(size=g["size"])
    orc = O.OracleScene(ft.SceneDescription(w.scene))
    img = orc.render(w.camera, g["size"], g["size"], w.epsilon, w.length, counters=True)
    assert int(bits(img["rgb"]).astype(np.uint64).sum()) == g["rgb_bits_sum"]
    assert int((img["leaf"].astype(np.int64) + 2).sum()) == g["leaf_sum"]
    assert img["counters"]["march_steps"] == g["march_steps"] and img["counters"]["shadow_rays"] == g["shadow_rays"]
    sel = np.array(g["pixels"])
    assert bits(img["rgb"][sel[:, 0], sel[:, 1]]).tolist() == g["pixel_rgb_bits"]
