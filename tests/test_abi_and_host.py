"""CPU-side tests of the boundary: the C-ABI library loads and exports every symbol the header declares,
the scene compiler's error behaviour mirrors the reference's, the host mirror flattens scenes correctly,
and no compute call works without a GPU (there is no fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene, scenes
from fraytracer_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = SdfForm.Primitive
GREY = SdfMaterial.createSolid((0.5, 0.5, 0.5))


def header_functions():
    src = open(os.path.join(ROOT, "include", "fraytracer_b200.h")).read()
    return sorted(set(re.findall(r"FT_API\s+[\w\s\*]+?\b(ft_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ft.load_library()
    declared = header_functions()
    assert declared and sorted(A.EXPORTS) == declared
    nm = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(ft_\w+)", nm))
    assert set(declared) <= exported
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ft_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", A.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_sizes_match_header():
    assert C.sizeof(A.ft_form_desc) == 52 and C.sizeof(A.ft_object_desc) == 20
    assert C.sizeof(A.ft_material_desc) == 16 and C.sizeof(A.ft_light_desc) == 28
    assert C.sizeof(A.ft_camera) == 48 and C.sizeof(A.ft_ray) == 32


def _compile_only(scene):
    return ft.SceneTracer(scene, device=A.FT_DEVICE_NONE)


def test_compile_only_handle_has_no_compute_path():
    w = scenes.default_scene(n_tori=30, size=16)
    tr = _compile_only(w.scene)
    info = tr.info()
    assert info["n_primitives"] == 32 and info["n_grids"] == 1 and info["device"] == -1
    for call in (lambda: tr.render(w.camera, w.size, w.epsilon, w.length),
                 lambda: tr.trace_rays(np.zeros((1, 8), np.float32) + 1),
                 lambda: tr.distance(np.zeros((1, 3), np.float32))):
        with pytest.raises(ft.FrayTracerError) as e:
            call()
        assert e.value.code == A.FT_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_scene_create_without_gpu_fails_loudly():
    n = C.c_int()
    lib = ft.load_library()
    lib.ft_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    w = scenes.default_scene(n_tori=10, size=8)
    with pytest.raises(ft.FrayTracerError) as e:
        ft.SceneTracer(w.scene, device=0)
    assert e.value.code == A.FT_ERR_NO_DEVICE


def _desc(obj, lights=()):
    return ft.SceneDescription(SdfScene(Object=obj, BackgroundColor=(0, 0, 0), Lights=list(lights)))


def _create(desc):
    lib = ft.load_library()
    h = C.c_void_p()
    st = lib.ft_scene_create(C.byref(desc.desc), A.FT_DEVICE_NONE, C.byref(h))
    msg = lib.ft_last_error().decode()
    if st == 0:
        lib.ft_scene_destroy(h)
    return st, msg


def test_compiler_error_behaviour_mirrors_reference():
    s = P.sphere(Center=(0, 0, 0), Radius=1.0)
    d = _desc(SdfObject.create(GREY, s))
    assert _create(d)[0] == A.FT_OK
    # empty union: failwith "No SdfObjects given." (SdfForm.fs:16)
    d._forms[0].kind = A.FT_FORM_UNION
    d._forms[0].child_count = 0
    st, msg = _create(d)
    assert st == A.FT_ERR_EMPTY_COMPOSITE and msg == "No SdfObjects given."
    d._forms[0].kind = A.FT_FORM_UNION_SMOOTH
    st, msg = _create(d)
    assert st == A.FT_ERR_EMPTY_COMPOSITE and msg == "blub"  # SdfForm.fs:71
    # a node kind the device cannot evaluate (user closure)
    d._forms[0].kind = 99
    st, msg = _create(d)
    assert st == A.FT_ERR_UNSUPPORTED_NODE and "closures" in msg
    d = _desc(SdfObject.create(GREY, s))
    d._materials[0].kind = 7
    assert _create(d)[0] == A.FT_ERR_UNSUPPORTED_NODE
    d = _desc(SdfObject.create(GREY, s), [SdfLight.point((0, 0, 0), (1, 1, 1))])
    d._lights[0].kind = 5
    assert _create(d)[0] == A.FT_ERR_UNSUPPORTED_NODE
    # malformed ids
    d = _desc(SdfObject.create(GREY, s))
    d._objects[0].form = 12
    assert _create(d)[0] == A.FT_ERR_INVALID_ARGUMENT
    d = _desc(SdfObject.create(GREY, s))
    d.desc.root_object = 3
    assert _create(d)[0] == A.FT_ERR_INVALID_ARGUMENT


def test_nesting_depth_limit():
    f = P.sphere(Center=(0, 0, 0), Radius=1.0)
    g = P.sphere(Center=(0.5, 0, 0), Radius=1.0)
    for _ in range(A.FT_MAX_DEPTH):
        f = SdfForm.subtract(f, g)
    assert _create(_desc(SdfObject.create(GREY, f)))[0] == A.FT_OK
    f = SdfForm.subtract(f, g)
    st, msg = _create(_desc(SdfObject.create(GREY, f)))
    assert st == A.FT_ERR_DEPTH


def test_description_flattening_and_leaf_ids():
    w = scenes.default_scene(n_tori=25, size=8)
    d = ft.SceneDescription(w.scene)
    assert d.desc.n_forms == 27 and d.desc.n_objects == 28 and d.desc.n_materials == 25 and d.desc.n_lights == 2
    ids = [d.leaf_id(o) for o in w.leaves]
    assert ids == list(range(25))  # generation order
    root = d._objects[d.desc.root_object]
    assert root.kind == A.FT_OBJECT_SUBTRACT
    # shared sub-objects are emitted once
    s = SdfObject.create(GREY, P.sphere(Center=(0, 0, 0), Radius=1.0))
    u = SdfObject.union([s, s, s])
    d = ft.SceneDescription(SdfScene(Object=u, BackgroundColor=(0, 0, 0), Lights=[]))
    assert d.desc.n_objects == 2 and d.desc.n_forms == 1 and d.desc.n_children == 3


def test_only_reified_nodes_cross_the_boundary():
    with pytest.raises(TypeError, match="FT_ERR_UNSUPPORTED_NODE"):
        SdfObject.create(GREY, lambda p: 1.0)
    with pytest.raises(TypeError, match="FT_ERR_UNSUPPORTED_NODE"):
        ft.Image.render(0.01, 30.0, ft.ImageSize(4, 4), scenes.default_camera(), lambda ray: (0, 0, 0))


def test_invalid_ray_parameters_are_rejected_before_any_launch():
    lib = ft.load_library()
    w = scenes.default_scene(n_tori=10, size=8)
    tr = _compile_only(w.scene)
    cam = w.camera.to_c()
    buf = np.zeros((8, 8, 3), np.float32)
    st = lib.ft_render(tr.handle, C.byref(cam), 8, 8, C.c_float(0.0), C.c_float(30.0), 0, buf.ctypes.data_as(C.c_void_p), None, None)
    assert st == A.FT_ERR_INVALID_ARGUMENT and b"epsilon" in lib.ft_last_error()
    st = lib.ft_render(tr.handle, C.byref(cam), 0, 8, C.c_float(0.01), C.c_float(30.0), 0, buf.ctypes.data_as(C.c_void_p), None, None)
    assert st == A.FT_ERR_INVALID_ARGUMENT


def test_strip_layout_matches_python_mirror():
    from fraytracer_b200 import distributed as D

    lib = ft.load_library()
    for (w_, h_, sw, world) in ((1920, 1080, 8, 1), (1920, 1080, 8, 8), (1001, 37, 8, 4), (7, 5, 8, 2), (100, 10, 16, 3)):
        spr, fl = C.c_int32(), C.c_int64()
        assert lib.ft_strip_layout(w_, h_, sw, world, C.byref(spr), C.byref(fl)) == 0
        assert (spr.value, fl.value) == D.strip_layout(w_, h_, sw, world)
        cols = sorted(x for r in range(world) for _, x0, n in D.owned_columns(w_, sw, r, world) for x in range(x0, x0 + n))
        assert cols == list(range(w_))  # every column owned exactly once


def test_every_compute_entry_point_refuses_a_compile_only_handle_and_bad_arguments():
    """no CPU fallback anywhere: the extension and multi-GPU entry points too answer FT_ERR_NO_DEVICE without a device,
    and validate their arguments before touching one"""
    w = scenes.default_scene(n_tori=30, size=16)
    tr = _compile_only(w.scene)
    cp = ft.CloudParams()
    calls = {
        "spectral": lambda: tr.render_spectral(w.camera, w.size, w.epsilon, w.length),
        "cloud": lambda: tr.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, 4),
        "gather": lambda: tr.render_gather_device(w.camera, w.size, w.epsilon, w.length, [0x1000]),
        "device": lambda: tr.render_device(w.camera, w.size, w.epsilon, w.length, 0x1000),
    }
    for name, call in calls.items():
        with pytest.raises(ft.FrayTracerError) as e:
            call()
        assert e.value.code == A.FT_ERR_NO_DEVICE, name
    bad = {
        "epsilon <= 0": lambda: tr.render_spectral(w.camera, w.size, 0.0, w.length),
        "wavelengths": lambda: tr.render_spectral(w.camera, w.size, w.epsilon, w.length, n_wavelengths=65),
        "bounces": lambda: tr.render_spectral(w.camera, w.size, w.epsilon, w.length, max_bounces=-1),
        "spp": lambda: tr.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, 0),
        "cloud step": lambda: tr.render_cloud(w.camera, w.size, w.epsilon, w.length, ft.CloudParams(step=0.0), 4),
        "perm range": lambda: tr.render_cloud(w.camera, w.size, w.epsilon, w.length,
                                              ft.CloudParams(perm=np.full(256, 300, np.int32)), 4),
        "no destination": lambda: tr.render_gather_device(w.camera, w.size, w.epsilon, w.length, []),
        "frame too large": lambda: tr.render(w.camera, ft.ImageSize(1 << 15, 1 << 15), w.epsilon, w.length,
                                             out=np.empty(3, np.float32)) if False else
        A.check(tr.lib.ft_render(tr.handle, C.byref(w.camera.to_c()), 1 << 15, 1 << 15, C.c_float(0.01), C.c_float(30.0), 0,
                                 C.c_void_p(0x1000), None, None)),
    }
    for name, call in bad.items():
        with pytest.raises(ft.FrayTracerError) as e:
            call()
        assert e.value.code == A.FT_ERR_INVALID_ARGUMENT, name
    tr.close()


def test_cloud_noise_tables_are_the_reference_generator():
    """CloudParams.default_tables regenerates Noise.gradientTable exactly as Noise.fs:31-33,72-76 draws it"""
    from fraytracer_b200.dotnet_random import DotNetRandom
    import math

    perm, grad = ft.CloudParams.default_tables()
    r = DotNetRandom(666)
    for _ in range(256):
        r.NextDouble()
    z = r.NextDouble() * 2.0 - 1.0
    rr = math.sqrt(1.0 - z * z)
    th = 2.0 * math.pi * r.NextDouble()
    assert grad[0].tolist() == [rr * math.cos(th), rr * math.sin(th), z]
    assert sorted(perm.tolist()) == list(range(256))
