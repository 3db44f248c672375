"""The C++ host mirror of the F# API (include/fraytracer_b200.hpp) and FrayTracer.Console restated over it
(examples/console.cpp = Program.fs): the reference is compiled code and its toolchain is absent, so the host side above the
C ABI exists in C++ as well as in Python (ctypes).  Both hosts must describe the same scene and render the same image."""
import json
import os
import subprocess

import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import _abi as A
from fraytracer_b200 import scenes

from conftest import bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "console")


@pytest.fixture(scope="module")
def console():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    return EXE


def test_cpp_console_describes_the_program_fs_scene_like_the_python_host(console):
    out = json.loads(subprocess.check_output([console, "--describe"]))
    tr = ft.SceneTracer(scenes.default_scene().scene, device=A.FT_DEVICE_NONE)
    info = tr.info()
    tr.close()
    for k in ("n_primitives", "n_grids", "n_cells", "n_entries", "max_depth"):
        assert out[k] == info[k], k
    # SURVEY's scratch statistics of the default scene's lookup: 14^3 cells, 152 538 candidate entries
    assert out["n_cells"] == 2744 and out["n_entries"] == 152538 and out["n_primitives"] == 1002
    small = json.loads(subprocess.check_output([console, "--describe", "--tori", "37"]))
    tr = ft.SceneTracer(scenes.default_scene(n_tori=37).scene, device=A.FT_DEVICE_NONE)
    assert small["n_entries"] == tr.info()["n_entries"] and small["n_primitives"] == 39
    tr.close()


def test_cpp_console_without_a_gpu_fails_loudly(console):
    n = __import__("ctypes").c_int()
    ft.load_library().ft_device_count(__import__("ctypes").byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([console, "--size", "32", "--tori", "20"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_console_renders_the_python_hosts_image_bit_for_bit(console, tmp_path):
    dump, bmp = str(tmp_path / "image.f32"), str(tmp_path / "result.bmp")
    out = subprocess.check_output([console, "--size", "320", "--dump", dump, "--out", bmp], text=True)
    assert "Rendering..." in out and "Time = " in out
    img = np.fromfile(dump, np.float32).reshape(320, 320, 3)
    w = scenes.default_scene(size=320)
    tr = ft.SceneTracer(w.scene, device=0)
    ref, _, _ = tr.render(w.camera, w.size, w.epsilon, w.length)
    assert np.array_equal(bits(img), bits(ref))
    want = ft.Image.toBitmap(2.2, 19, ref, device=0)
    tr.close()
    raw = open(bmp, "rb").read()
    assert raw[:2] == b"BM" and len(raw) >= 54 + 320 * 320 * 3
    tmp = str(tmp_path / "python.bmp")
    ft.Image.saveBitmap(tmp, want)
    assert raw == open(tmp, "rb").read()
