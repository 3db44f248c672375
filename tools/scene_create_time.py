"""ft_scene_create wall time (host compile + upload + sub-grid build) of the reference-defined configs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A

for make in (scenes.default_scene, scenes.spheres_capsules_scene, scenes.mixed_scene):
    w = make()
    t0 = time.perf_counter(); d = ft.SceneDescription(w.scene); t_desc = time.perf_counter() - t0
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        tr = ft.SceneTracer(w.scene, device=0)
        ts.append(time.perf_counter() - t0)
        sg = tr.subgrid_info(); info = tr.info()
        tr.close()
    t0 = time.perf_counter(); tr = ft.SceneTracer(w.scene, device=A.FT_DEVICE_NONE); t_host = time.perf_counter() - t0
    print(f"{w.name:34s} python description {t_desc*1e3:6.1f} ms | SceneTracer total first {ts[0]*1e3:7.1f} ms, then {min(ts[1:])*1e3:6.1f} ms "
          f"(compile-only handle {t_host*1e3:6.1f} ms; sub-grid build {sg['build_ms']:.2f} ms, {sg['device_bytes']/1e6:.1f} MB; grids on gpu {info['grids_built_on_gpu']})")
