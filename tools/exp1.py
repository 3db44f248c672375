"""experiments: band count, register caps, fast-mode error distribution"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import fraytracer_b200 as ft
    from fraytracer_b200 import scenes, _abi as A
    for w in (scenes.default_scene(), scenes.spheres_capsules_scene()):
        tr = ft.SceneTracer(w.scene, 0)
        best = 1e9
        for it in range(5):
            _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length)
            if it: best = min(best, st["kernel_ms"])
        print(f"   {w.name}: kernel {best:.3f} ms", flush=True)
    sys.exit(0)
for lib in ("", "build/lib_minb6.so", "build/lib_minb8.so"):
    for bands in (1, 2, 6, 16):
        env = dict(os.environ, FT_BANDS=str(bands))
        if lib: env["FT_LIB"] = os.path.abspath(lib)
        print(f"lib={lib or 'default(96 regs)'} bands={bands}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
for w in (scenes.default_scene(), scenes.spheres_capsules_scene()):
    tr = ft.SceneTracer(w.scene, 0)
    a, la, _ = tr.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True)
    b, lb, _ = tr.render(w.camera, w.size, w.epsilon, w.length, want_leaf=True, flags=A.FT_FLAG_FAST_MATH)
    same = la == lb
    rel = (np.abs(a - b) / np.maximum(np.abs(a), 1e-3)).max(-1)[same & (la >= 0)]
    print(w.name, "leaf mismatch", 1 - same.mean(), "rel err percentiles 50/90/99/99.9/max",
          [float(np.percentile(rel, q)) for q in (50, 90, 99, 99.9, 100)], "frac>1e-4", (rel > 1e-4).mean(), "frac>1e-2", (rel > 1e-2).mean())
