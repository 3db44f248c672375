"""experiment (1 GPU): cost of the strip decomposition -- rank r's share of an 8K frame for world = 8 and several strip widths"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from fraytracer_b200.distributed import strip_layout
w = scenes.mixed_scene(width=7680, height=4320)
tr = ft.SceneTracer(w.scene, 0)
X, Y = w.size.X, w.size.Y
stream = torch.cuda.current_stream().cuda_stream
def timed(fn, n=4):
    best = 1e9
    for it in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it: best = min(best, e0.elapsed_time(e1))
    return best
full = torch.empty((X, Y, 3), dtype=torch.float32, device="cuda")
t1 = timed(lambda: tr.render_device(w.camera, w.size, w.epsilon, w.length, full.data_ptr(), stream=stream))
print(f"whole frame, world 1: {t1:.3f} ms  (1/8 = {t1/8:.3f})", flush=True)
for world in (8,):
    for sw in (4, 8, 16, 32, 64, 128, 960):
        spr, nfl = strip_layout(X, Y, sw, world)
        buf = torch.empty(nfl, dtype=torch.float32, device="cuda")
        ts = [timed(lambda r=r: tr.render_device(w.camera, w.size, w.epsilon, w.length, buf.data_ptr(), stream=stream,
                                                   strip_width=sw, rank=r, world=world)) for r in (0, 3, 7)]
        print(f"world {world} strip_width {sw}: ranks 0/3/7 {ts[0]:.3f} {ts[1]:.3f} {ts[2]:.3f} ms", flush=True)
