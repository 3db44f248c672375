"""experiment: extension kernels (spectral config 3 at 4K, cloud config 4 at 1080p / 16 spp) for several library builds"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import fraytracer_b200 as ft
    from fraytracer_b200 import scenes
    w = scenes.spectral_glass_scene()
    tr = ft.SceneTracer(w.scene, 0)
    best = 1e9
    for it in range(3):
        _, st = tr.render_spectral(w.camera, w.size, w.epsilon, w.length, 16, 8, 19)
        if it: best = min(best, st["kernel_ms"])
    print(f"   spectral 4K: {best:.1f} ms", flush=True)
    w = scenes.cloud_scene()
    tr = ft.SceneTracer(w.scene, 0)
    cp = ft.CloudParams()
    best = 1e9
    for it in range(3):
        _, st = tr.render_cloud(w.camera, w.size, w.epsilon, w.length, cp, 16, 19)
        if it: best = min(best, st["kernel_ms"])
    print(f"   cloud 1080p 16spp: {best:.1f} ms", flush=True)
    sys.exit(0)
for lib in sys.argv[1:] or ["default"]:
    env = dict(os.environ)
    if lib != "default": env["FT_LIB"] = os.path.abspath(lib)
    print(f"lib={lib}", flush=True)
    subprocess.run([sys.executable, __file__, "child"], env=env)
