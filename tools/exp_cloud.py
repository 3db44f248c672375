"""config 4 (EXTENSION, no reference implementation): time ft_render_cloud at 1080p, 64 samples per pixel."""
import json
import sys
import time

sys.path.insert(0, ".")
import fraytracer_b200 as ft
from fraytracer_b200 import scenes

W, H, SPP = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 64)
w = scenes.cloud_scene(width=W, height=H)
tr = ft.SceneTracer(w.scene, device=0)
cp = ft.CloudParams()
res = []
for spp in (1, SPP, SPP):
    t = time.time()
    img, st = tr.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, spp, 19)
    wall = time.time() - t
    rays = st["primary_rays"] + st["shadow_rays"]
    res.append(dict(spp=spp, kernel_ms=st["kernel_ms"], total_ms=st["total_ms"], wall_ms=wall * 1e3, rays=rays,
                    mrays_s=rays / st["kernel_ms"] / 1e3, track_steps=st["march_steps"], scattered=st["reserved"],
                    msamples_s=st["primary_rays"] / st["kernel_ms"] / 1e3, escaped=st["escaped_rays"], capped=st["capped_rays"]))
print(json.dumps(dict(workload=w.name, image=f"{W}x{H}", runs=res, mean_rgb=[float(x) for x in img.mean(axis=(0, 1))])))
