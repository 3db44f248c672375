"""torchrun check (N >= 2 GPUs): the fused render + P2P gather framebuffer equals the single-GPU render bit for bit on every
rank, for both gather modes, several frames in a row (frame-parity buffers).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_p2p.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from fraytracer_b200.distributed import DistributedRenderer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for w, (X, Y) in ((scenes.default_scene(n_tori=300), (333, 251)), (scenes.spheres_capsules_scene(), (640, 360))):
    size = ft.ImageSize(X, Y)
    tr = ft.SceneTracer(w.scene, device=local)
    ref, _, _ = tr.render(w.camera, size, w.epsilon, w.length)
    for mode, sw in (("p2p", None), ("p2p", 8), ("nccl", None)):
        r = DistributedRenderer(tr, size, strip_width=sw, gather=mode)
        for frame in range(3):
            img = r.render(w.camera, w.epsilon, w.length).cpu().numpy()
            same = np.array_equal(img.view(np.uint32), ref.view(np.uint32))
            ok &= same
            print(f"rank {rank}/{world} {w.name} {X}x{Y} {mode} strip {r.strip_width} frame {frame}: identical={same}", flush=True)
        if mode == "p2p":  # shared host framebuffer: every rank's kernel stores its strips over its own PCIe link
            for frame in range(2):
                host = r.render_to_host(w.camera, w.epsilon, w.length)
                same = np.array_equal(host.view(np.uint32), ref.view(np.uint32))
                ok &= same
                print(f"rank {rank}/{world} {w.name} {X}x{Y} host frame {frame}: identical={same}", flush=True)
        r.close()
    tr.close()
# the extension paths (no reference implementation): N-rank frame == 1-rank frame
for kind, w, kw in (("spectral", scenes.spectral_glass_scene(width=320, height=180), dict(n_wavelengths=8, max_bounces=6, seed=5)),
                    ("cloud", scenes.cloud_scene(width=320, height=180), dict(samples_per_pixel=6, seed=5))):
    tr = ft.SceneTracer(w.scene, device=local)
    if kind == "spectral":
        ref, _ = tr.render_spectral(w.camera, w.size, w.epsilon, w.length, 8, 6, 5)
    else:
        ref, _ = tr.render_cloud(w.camera, w.size, w.epsilon, w.length, ft.CloudParams(), 6, 5)
    r = DistributedRenderer(tr, w.size, gather="p2p")
    for frame in range(2):
        img, _ = r.render_extension(kind, w.camera, w.epsilon, w.length, **kw)
        same = np.array_equal(img.cpu().numpy().view(np.uint32), ref.view(np.uint32))
        ok &= same
        print(f"rank {rank}/{world} {w.name} {kind} strip {r.strip_width} frame {frame}: identical={same}", flush=True)
    r.close()
    tr.close()
t = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(int(t.item()) != 0)
