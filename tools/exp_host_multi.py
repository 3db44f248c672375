"""torchrun experiment: end to end into ONE shared, page-locked host frame from N GPUs (DistributedRenderer.render_to_host), config 2
weak-scaling frame (or FT_EXP_WORKLOAD=config5: one 8K frame).  Environment switches of the library select the tile order."""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import fraytracer_b200 as ft
from fraytracer_b200.distributed import DistributedRenderer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = os.environ.get("FT_EXP_WORKLOAD", "config2")
w = bench.workload(which, world, which == "config5")
tr = ft.SceneTracer(w.scene, device=local)
sw = int(os.environ.get("FT_EXP_STRIP", "0"))
if sw < 0:  # one contiguous slab per rank
    sw = -(-(-(-w.size.X // world)) // 4) * 4
r = DistributedRenderer(tr, w.size, strip_width=sw or None, gather="p2p")
r.host_frame()
ts = []
for it in range(16):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    r.render_to_host(w.camera, w.epsilon, w.length, 0)
    ts.append(time.perf_counter() - t0)
t = torch.tensor(sorted(ts[4:])[len(ts[4:]) // 2], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    X, Y = w.size.X, w.size.Y
    print(json.dumps({"n_gpus": world, "strip_width": r.strip_width, "workload": w.name, "image": f"{X}x{Y}", "e2e_ms_median": 1e3 * float(t),
                      "GB_per_s": X * Y * 12 / float(t) / 1e9,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("FT_") and k != "FT_EXP_WORKLOAD"}}), flush=True)
r.close(); tr.close()
dist.destroy_process_group()
