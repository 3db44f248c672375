"""Render one config a few times (for ncu captures).  usage: prof_render.py <config1|config2|config5> [reps] [flags]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fraytracer_b200 as ft
from fraytracer_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "config1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = {"config1": scenes.default_scene, "config2": scenes.spheres_capsules_scene,
     "config5": lambda: scenes.mixed_scene(width=3840, height=2160)}[name]()
tr = ft.SceneTracer(w.scene, 0)
for _ in range(reps):
    _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length, flags=flags)
print(name, st)
