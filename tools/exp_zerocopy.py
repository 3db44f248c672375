"""experiment: ft_render into pinned+mapped host memory -- kernel stores straight to host (zero-copy) vs render + band copies"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
lib = ft.load_library()
for w in (scenes.spheres_capsules_scene(), scenes.default_scene(), scenes.mixed_scene(width=3840, height=2160)):
    tr = ft.SceneTracer(w.scene, 0)
    X, Y = w.size.X, w.size.Y
    host = C.c_void_p(); A.check(lib.ft_host_alloc(C.byref(host), X * Y * 12))
    out = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3))
    ref = None
    for mode in ("copy", "zerocopy"):
        if mode == "copy": os.environ["FT_NO_ZEROCOPY"] = "1"
        else: os.environ.pop("FT_NO_ZEROCOPY", None)
        best = 1e9
        for it in range(10):
            out[:] = 0
            t0 = time.perf_counter()
            _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, out=out)
            dt = time.perf_counter() - t0
            if it > 1: best = min(best, dt)
        if ref is None: ref = out.copy()
        same = np.array_equal(ref.view(np.uint32), out.view(np.uint32))
        print(f"{w.name} {mode}: wall {best*1e3:.3f} ms (device total {st['total_ms']:.3f}, kernels {st['kernel_ms']:.3f}, launches {st['launches']}) identical={same}", flush=True)
    lib.ft_host_free(host)
