"""experiment: register caps with the current kernel (single launch)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import fraytracer_b200 as ft
    from fraytracer_b200 import scenes
    for w in (scenes.default_scene(), scenes.spheres_capsules_scene(), scenes.mixed_scene(width=3840, height=2160)):
        tr = ft.SceneTracer(w.scene, 0)
        best = 1e9
        for it in range(5):
            _, _, st = tr.render(w.camera, w.size, w.epsilon, w.length)
            if it: best = min(best, st["kernel_ms"])
        rays = st["primary_rays"] + st["shadow_rays"]
        print(f"   {w.name}: kernel {best:.3f} ms  {rays/best/1e3:.0f} Mrays/s", flush=True)
    sys.exit(0)
libs = sys.argv[1:] or [""]
for lib in libs:
    env = dict(os.environ, FT_BANDS="1")
    if lib and lib != "default": env["FT_LIB"] = os.path.abspath(lib)
    print(f"lib={lib}", flush=True)
    subprocess.run([sys.executable, __file__, "child"], env=env)
