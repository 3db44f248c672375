import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FT_PRINT_LONG"] = "1"
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
for w in (scenes.default_scene(), scenes.spheres_capsules_scene(), scenes.mixed_scene(width=3840, height=2160)):
    tr = ft.SceneTracer(w.scene, 0)
    _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, flags=A.FT_FLAG_COUNTERS)
    print(w.name, {k: st[k] for k in ("primary_rays", "primary_hits", "shadow_rays", "escaped_rays", "march_steps", "eval_flops")}, flush=True)
