"""config 3 (EXTENSION, no reference implementation): time ft_render_spectral at 4K, 16 wavelengths per pixel."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import fraytracer_b200 as ft
from fraytracer_b200 import scenes

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
w = scenes.spectral_glass_scene(width=W, height=H)
tr = ft.SceneTracer(w.scene, device=0)
res = []
for mb in (8, 0, 1, 2, 4, 8, 8):
    t = time.time()
    img, st = tr.render_spectral(w.camera, w.size, w.epsilon, w.length, 16, mb, 19)
    wall = time.time() - t
    rays = st["primary_rays"] + st["shadow_rays"]
    res.append(dict(max_bounces=mb, kernel_ms=st["kernel_ms"], total_ms=st["total_ms"], wall_ms=wall * 1e3, rays=rays,
                    mrays_s=rays / st["kernel_ms"] / 1e3, glass_pixels=st["reserved"], launches=st["launches"],
                    escaped=st["escaped_rays"], capped=st["capped_rays"]))
print(json.dumps(dict(workload=w.name, image=f"{W}x{H}", runs=res, mean_rgb=[float(x) for x in img.mean(axis=(0, 1))])))
