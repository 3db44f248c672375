"""The oracle against a second, independent restatement of the WHOLE ray-test (tests/py_trace.py: plain Python over numpy
float32 scalars, written from the F# source, no code shared with oracle/ft_oracle.c) on BASELINE config 1 -- Program.fs's
1000 tori, clip sphere, cut sphere, both lights -- bit for bit.  Two readers of the reference's source can still share a
misreading, and neither is the reference itself (parity stays unpinned); what this removes is the
"one file by one author" risk the compiler-vs-oracle tests cannot catch."""
import numpy as np

import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from oracle import oracle as O

import py_trace as T
from conftest import bits


def _config1(n_tori=1000):
    b = scenes.SceneBuilder(19)
    tori, colors = [], []
    for _ in range(n_tori):
        o = b.randomTorus()
        p = o.form.params
        tori.append((p[0:3], p[3:6], p[6], p[7]))
        colors.append(o.material.rgb)
    lights = [("dir", (-0.5, -1.0, 1.0), (0.5, 0.5, 0.5)), ("point", (-0.5, 0.0, -2.0), (10.0, 0.0, 0.0))]  # Program.fs:79-82
    return T.Scene(tori, colors, ((0.0, 0.0, 0.0), 3.5), ((-0.5, 1.0, -2.0), 2.5), lights, (0.1, 0.1, 0.1))


def test_whole_trace_config1_32x32_bit_exact():
    size = 32
    w = scenes.default_scene(size=size)
    desc = ft.SceneDescription(w.scene)
    ref = O.OracleScene(desc).render(O.camera_look_at((0, 0, -10), (0, 0, 0), (0, 1, 0), O.lens_create(60.0)), size, size,
                                     w.epsilon, w.length)
    scene = _config1()
    cam = T.look_at((0, 0, -10), (0, 0, 0), (0, 1, 0), 60.0)
    got = T.render(scene, *cam, size, size, 0.01, 30.0)
    leaves = {desc.leaf_id(o): i for i, o in enumerate(w.leaves)}
    n_hit = 0
    for (x, y), (rgb, leaf) in got.items():
        want_leaf = int(ref["leaf"][x, y])
        assert (leaf if leaf < 0 else leaf) == (-1 if want_leaf < 0 else leaves[want_leaf]), (x, y)
        assert bits(np.array(rgb, np.float32)).tolist() == bits(ref["rgb"][x, y]).tolist(), (x, y)
        n_hit += leaf >= 0
    assert 200 < n_hit < size * size  # the frame really shows the scene


def test_whole_trace_full_resolution_pixels_bit_exact():
    """a sample of pixels of the real 1000 x 1000 frame (silhouettes included: every 37th pixel of three columns)"""
    w = scenes.default_scene()
    desc = ft.SceneDescription(w.scene)
    X = Y = 1000
    orc = O.OracleScene(desc)
    cam_c = O.camera_look_at((0, 0, -10), (0, 0, 0), (0, 1, 0), O.lens_create(60.0))
    scene = _config1()
    cam = T.look_at((0, 0, -10), (0, 0, 0), (0, 1, 0), 60.0)
    leaves = {desc.leaf_id(o): i for i, o in enumerate(w.leaves)}
    for x in (333, 500, 761):
        ref = orc.render(cam_c, X, Y, w.epsilon, w.length, x_begin=x, x_end=x + 1)
        pix = [(x, y) for y in range(5, Y, 37)]
        got = T.render(scene, *cam, X, Y, 0.01, 30.0, pixels=pix)
        for (px, py), (rgb, leaf) in got.items():
            want_leaf = int(ref["leaf"][px, py])
            assert leaf == (-1 if want_leaf < 0 else leaves[want_leaf]), (px, py)
            assert bits(np.array(rgb, np.float32)).tolist() == bits(ref["rgb"][px, py]).tolist(), (px, py)
