"""Development probe (GPU box): time the render kernels of the three reference-defined configs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
import ctypes as C

lib = ft.load_library()
tf, mhz = C.c_float(), C.c_float()
A.check(lib.ft_measure_fp32_peak(0, C.byref(tf), C.byref(mhz)))
print(f"fp32 FFMA peak {tf.value:.1f} TFLOP/s (~{mhz.value:.0f} MHz)")
cfgs = [scenes.default_scene(), scenes.spheres_capsules_scene(), scenes.mixed_scene(width=3840, height=2160)]
only = sys.argv[1:] 
for w in cfgs:
    if only and not any(o in w.name for o in only): continue
    tr = ft.SceneTracer(w.scene, 0)
    X, Y = w.size.X, w.size.Y
    for flags, name in ((0, "strict"), (A.FT_FLAG_FAST_MATH, "fast")):
        best = None
        for it in range(4):
            rgb, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, flags=flags)
            if it and (best is None or st["kernel_ms"] < best["kernel_ms"]): best = st
        rays = best["primary_rays"] + best["shadow_rays"]
        print(f"{w.name} {X}x{Y} {name}: kernel {best['kernel_ms']:.3f} ms total {best['total_ms']:.3f} ms  "
              f"rays {rays} -> {rays/best['kernel_ms']/1e3:.1f} Mrays/s  hits {best['primary_hits']} capped {best['capped_rays']}")
    _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, flags=A.FT_FLAG_COUNTERS)
    rays = st["primary_rays"] + st["shadow_rays"]
    print(f"   counters: steps {st['march_steps']} ({st['march_steps']/rays:.1f}/ray) flops {st['eval_flops']:.3e} "
          f"({st['eval_flops']/max(st['march_steps'],1):.0f}/step) kernel(counting) {st['kernel_ms']:.3f} ms")
    tr.close()
