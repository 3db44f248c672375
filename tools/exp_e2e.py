"""experiment: wall time of ft_render (host pinned output) against the number of column bands"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
lib = ft.load_library()
for w in (scenes.spheres_capsules_scene(), scenes.default_scene()):
    tr = ft.SceneTracer(w.scene, 0)
    X, Y = w.size.X, w.size.Y
    host = C.c_void_p(); A.check(lib.ft_host_alloc(C.byref(host), X * Y * 12))
    out = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3))
    for bands in (1, 2, 3, 4, 6, 8, 12):
        os.environ["FT_BANDS"] = str(bands)
        best = 1e9
        for it in range(8):
            t0 = time.perf_counter()
            _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, out=out)
            dt = time.perf_counter() - t0
            if it > 1: best = min(best, dt)
        print(f"{w.name} bands={bands}: wall {best*1e3:.3f} ms  (device total {st['total_ms']:.3f}, kernels {st['kernel_ms']:.3f})", flush=True)
    lib.ft_host_free(host)
