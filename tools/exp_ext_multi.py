"""torchrun experiment: the extension workloads (config 3 at 4K / 16 wavelengths, config 4 at 1080p / 64 spp) cut over N GPUs
(strong scaling of one frame; P2P gather).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from fraytracer_b200.distributed import DistributedRenderer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"n_gpus": world}
for kind, w, kw in (("spectral", scenes.spectral_glass_scene(), dict(n_wavelengths=16, max_bounces=8, seed=19)),
                    ("cloud", scenes.cloud_scene(), dict(samples_per_pixel=64, seed=19))):
    tr = ft.SceneTracer(w.scene, device=local)
    r = DistributedRenderer(tr, w.size, gather="p2p")
    best = 1e9
    rays = 0
    for it in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        _, st = r.render_extension(kind, w.camera, w.epsilon, w.length, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if it:
            best = min(best, dt)
    t = torch.tensor([float(st["primary_rays"] + st["shadow_rays"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    rays = float(t.item())
    out[kind] = {"workload": w.name, "ms_per_frame": best * 1e3, "rays": rays, "mrays_s": rays / best / 1e6}
    r.close()
    tr.close()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
