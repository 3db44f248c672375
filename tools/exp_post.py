"""timing of the post-process kernels and of scene compilation (development probe)"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
lib = ft.load_library()
for make in (scenes.default_scene, scenes.spheres_capsules_scene, scenes.mixed_scene):
    w = make()
    t0 = time.perf_counter(); d = ft.SceneDescription(w.scene); t1 = time.perf_counter()
    tr = ft.SceneTracer(w.scene, 0); t2 = time.perf_counter()
    print(f"{w.name}: flatten {1e3*(t1-t0):.1f} ms, ft_scene_create (flatten+compile+upload) {1e3*(t2-t1):.1f} ms", tr.info()["n_entries"])
X, Y = 1920, 1080
img = torch.rand((X, Y, 3), device="cuda") * 2
mx = torch.zeros(1, device="cuda"); out = torch.zeros((Y, X, 3), dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(5):
    ev[0].record()
    A.check(lib.ft_image_max_device(C.c_void_p(img.data_ptr()), img.numel(), C.c_void_p(mx.data_ptr()), None))
    ev[1].record()
    A.check(lib.ft_tonemap_device(C.c_void_p(img.data_ptr()), X, Y, C.c_float(2.2), C.c_void_p(mx.data_ptr()), C.c_float(0), C.c_uint64(19), C.c_void_p(out.data_ptr()), 3 * X, None))
    ev[2].record(); torch.cuda.synchronize()
print(f"1080p: image max {ev[0].elapsed_time(ev[1])*1e3:.1f} us, tonemap+pack {ev[1].elapsed_time(ev[2])*1e3:.1f} us")
