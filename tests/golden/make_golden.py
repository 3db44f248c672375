"""Generates tests/golden/oracle_config1_64.json from the CPU oracle itself (regression fixture: it pins the
oracle against accidental change; it is NOT reference output -- no .NET runtime exists here)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from oracle import oracle as O

size = 64
w = scenes.default_scene(size=size)
orc = O.OracleScene(ft.SceneDescription(w.scene))
img = orc.render(w.camera, size, size, w.epsilon, w.length, counters=True)
b = np.ascontiguousarray(img["rgb"]).view(np.uint32)
pix = [[x, y] for x in range(4, size, 12) for y in range(5, size, 12)]
sel = np.array(pix)
out = dict(size=size, rgb_bits_sum=int(b.astype(np.uint64).sum()), leaf_sum=int((img["leaf"].astype(np.int64) + 2).sum()),
           march_steps=img["counters"]["march_steps"], shadow_rays=img["counters"]["shadow_rays"],
           pixels=pix, pixel_rgb_bits=b[sel[:, 0], sel[:, 1]].tolist())
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_config1_64.json"), "w"))
print("written", out["rgb_bits_sum"], out["march_steps"])
