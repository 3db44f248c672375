"""world_size-2 gloo test of the multi-GPU path's host logic: strip ownership,
compact per-rank buffers, all-gather, de-interleave.  The pixels come from the oracle here (this is a CPU
test of the partitioning, not of the kernels); the N-rank framebuffer must equal the 1-rank one bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, X, Y, sw, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fraytracer_b200 as ft
    from fraytracer_b200 import distributed as D, scenes
    from oracle import oracle as O

    w = scenes.default_scene(n_tori=80, size=X)
    orc = O.OracleScene(ft.SceneDescription(w.scene))
    spr, floats = D.strip_layout(X, Y, sw, world)
    local = torch.zeros((spr * sw, Y, 3), dtype=torch.float32)
    for lc, x0, n in D.owned_columns(X, sw, rank, world):
        img = orc.render(w.camera, X, Y, w.epsilon, w.length, threads=2, x_begin=x0, x_end=x0 + n)
        local[lc:lc + n] = torch.from_numpy(img["rgb"][x0:x0 + n])
    assert local.numel() == floats
    gathered = torch.empty((world, spr * sw, Y, 3), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(-1, Y, 3), local)
    final = D.deinterleave_reference(gathered, X, Y, sw, world)
    if rank == 0:
        full = orc.render(w.camera, X, Y, w.epsilon, w.length, threads=2)["rgb"]
        ok = np.array_equal(final.numpy().view(np.uint32), full.view(np.uint32))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("X,Y,sw", [(44, 20, 8), (37, 11, 4)])
def test_two_rank_strip_gather_equals_single_rank(tmp_path, X, Y, sw):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() + X) % 2000
    mp.spawn(_worker, args=(2, port, X, Y, sw, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_strip_partition_covers_every_column_exactly_once():
    """host logic of the multi-GPU partition: for any width / world / strip width the ranks' strips tile the frame"""
    from fraytracer_b200.distributed import auto_strip_width, owned_columns, strip_layout

    for width in (1, 7, 8, 333, 1920, 2712, 5432, 7680):
        for world in (1, 2, 3, 4, 8):
            for sw in (4, 8, 20, 32, auto_strip_width(width, world)):
                seen = [0] * width
                for rank in range(world):
                    spr, floats = strip_layout(width, 10, sw, world)
                    cols = owned_columns(width, sw, rank, world)
                    assert len(cols) <= spr
                    for local, x0, n in cols:
                        assert local % sw == 0 and 0 < n <= sw
                        for x in range(x0, x0 + n):
                            seen[x] += 1
                    assert floats == (width if world == 1 else spr * sw) * 10 * 3
                assert seen == [1] * width, (width, world, sw)
    for width, world in ((7680, 8), (5432, 8), (2712, 2), (640, 8), (100, 16)):
        sw = auto_strip_width(width, world)
        assert sw % 4 == 0 and 8 <= sw <= 32
