"""scene-compile time, host vs device grid build, on a scene with many small primitives (development probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import SdfForm, SdfMaterial, SdfObject, SdfScene, scenes
P = SdfForm.Primitive
tr = ft.SceneTracer(scenes.default_scene(n_tori=10).scene, 0)  # CUDA context
for n, rs in ((4096, 0.25), (16384, 0.12), (32768, 0.1)):
    rng = np.random.default_rng(1)
    m = SdfMaterial.createSolid((0.5, 0.5, 0.5))
    c = rng.normal(size=(n, 3)); c = c / np.linalg.norm(c, axis=1, keepdims=True) * rng.uniform(0, 4, size=(n, 1)) ** (1 / 3) * 1.6
    objs = [SdfObject.create(m, P.sphere(Center=c[i], Radius=float(rng.uniform(0.3, 1.0) * rs))) for i in range(n)]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=(0.1, 0.1, 0.1), Lights=scenes.default_lights())
    d = ft.SceneDescription(scene)
    for mode in ("cpu", "gpu"):
        os.environ["FT_GRID_BUILD"] = mode
        t0 = time.perf_counter(); t = ft.SceneTracer(scene, 0); dt = time.perf_counter() - t0
        i = t.info()
        print(f"n={n} radius x{rs}: {mode} ft_scene_create {dt*1e3:8.1f} ms  cells {i['n_cells']} entries {i['n_entries']} gpu grids {i['grids_built_on_gpu']}", flush=True)
        t.close()
