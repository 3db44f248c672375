"""Multi-GPU image partitioning: pixels are independent, the compiled scene is small
and replicated on every GPU, the image is cut into column strips dealt round-robin to the ranks
(strip s -> rank s % world), every rank renders its strips into a compact buffer, one NCCL all-gather
moves the framebuffer over NVLink and a de-interleave kernel restores the x-major FColor[X,Y]  (gather="nccl").

gather="p2p" (the default with more than one rank) fuses the gather into the render kernel: every rank maps its peers'
framebuffers through CUDA IPC and the kernel stores each finished pixel at its final position in all of them -- the
transfer rides over NVLink / NVSwitch pixel by pixel while the rest of the frame is still being traced; a 4-byte NCCL
all-reduce afterwards is the only collective (it tells every rank that all writers are done).

One process per GPU; torch.distributed is only the plumbing around the C ABI (`ft_render_device`,
`ft_deinterleave_device`).  The N-GPU framebuffer must be bit-identical to the 1-GPU one.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi as A


def auto_strip_width(width: int, world: int) -> int:
    """Strip width for `world` ranks, a multiple of the 4-pixel tile width.  Cost per pixel varies > 100x over a frame, so
    every rank should own many strips spread over it; wider strips keep the tiles a GPU has in flight closer together
    (cells and candidate lists shared in L1/L2).  Measured on B200 (8K frame, 8 ranks, slowest rank): 8 columns 4.0 ms,
    32 columns 3.95 ms, 120 columns 4.22 ms (imbalance), contiguous slabs 2.3 ... 4.2 ms -- a shallow optimum, so about
    32 strips per rank, at most 32 columns."""
    if world <= 1:
        return 8
    return max(8, min(32, (width // (world * 32)) // 4 * 4))


def strip_layout(width: int, height: int, strip_width: int, world: int):
    """(strips_per_rank, floats_per_rank) of the compact per-rank buffer -- same arithmetic as ft_strip_layout"""
    strips = (width + strip_width - 1) // strip_width
    spr = (strips + world - 1) // world
    cols = width if world == 1 else spr * strip_width
    return spr, cols * height * 3


def owned_columns(width: int, strip_width: int, rank: int, world: int):
    """[(local_column_begin, image_column_begin, n_columns)] of the strips rank owns, in buffer order"""
    strips = (width + strip_width - 1) // strip_width
    out = []
    for local, s in enumerate(range(rank, strips, world)):
        x0 = s * strip_width
        out.append((local * strip_width, x0, min(strip_width, width - x0)))
    return out


def deinterleave_reference(gathered, width: int, height: int, strip_width: int, world: int):
    """numpy/torch restatement of ft_deinterleave_device for the CPU (gloo) tests:
    gathered [world, spr*strip_width, height, C] -> [width, height, C]"""
    out = gathered.new_zeros((width, height, gathered.shape[-1])) if hasattr(gathered, "new_zeros") else None
    if out is None:
        import numpy as np

        out = np.zeros((width, height, gathered.shape[-1]), gathered.dtype)
    for rank in range(world):
        for lc, x0, n in owned_columns(width, strip_width, rank, world):
            out[x0:x0 + n] = gathered[rank, lc:lc + n]
    return out


class _DeviceArray:
    """raw device pointer -> __cuda_array_interface__ (float32), so that torch can view library-allocated memory"""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _interleave_numa(buf, nbytes: int) -> bool:
    """Spread the pages of the shared host framebuffer over all NUMA nodes (mbind MPOL_INTERLEAVE on the tmpfs mapping: a
    shared policy, valid for every process that maps it).  With N GPUs behind two sockets writing one frame, first-touch
    placement would put every page on the node of rank 0 and send half of the GPUs' stores across the socket interconnect
    into one node's memory controllers.  Best effort: False (and first-touch placement) when the kernel refuses.
    FT_HOST_INTERLEAVE=0 disables it."""
    import ctypes
    import glob

    if os.environ.get("FT_HOST_INTERLEAVE", "1") == "0":
        return False
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    if len(nodes) < 2:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(buf))
        page = os.sysconf("SC_PAGE_SIZE")
        start = addr & ~(page - 1)
        length = ((addr + nbytes + page - 1) & ~(page - 1)) - start
        mask = (ctypes.c_ulong * 16)()
        for n in nodes:
            mask[n // 64] |= 1 << (n % 64)
        MPOL_INTERLEAVE, SYS_mbind = 3, 237  # x86_64
        rc = libc.syscall(SYS_mbind, ctypes.c_void_p(start), ctypes.c_ulong(length), ctypes.c_int(MPOL_INTERLEAVE), mask,
                          ctypes.c_ulong(16 * 64), ctypes.c_uint(0))
        return rc == 0
    except Exception:  # noqa: BLE001
        return False


class DistributedRenderer:
    """Renders one frame cooperatively on all ranks of `group`; every rank ends with the full framebuffer.

    A rank may render its strips in `chunks` launches, the NCCL all-gather of chunk k being issued asynchronously
    behind chunk k+1's kernel.  Measured on 2 x B200 this does NOT pay: the persistent render grid fills every SM
    slot, so the collective's kernel cannot start before the next chunk drains, and every chunk adds a tail
    (3.09 ms per step with 2 chunks against 2.60 ms with 1).  The default is therefore one chunk: render, one
    all-gather, de-interleave."""

    def __init__(self, tracer, size, strip_width: int | None = None, group=None, chunks: int = 1, gather: str | None = None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.tracer, self.size, self.group = tracer, size, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.strip_width = strip_width = int(strip_width) if strip_width else auto_strip_width(size.X, self.world)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        dev = torch.device("cuda", tracer.device)
        X, Y = size.X, size.Y
        self.spr, _ = strip_layout(X, Y, strip_width, self.world)
        self.gather = gather or ("p2p" if self.world > 1 else "nccl")
        self.chunk_ranges = []
        self._owned, self._imported = [], []
        if self.world > 1 and self.gather == "p2p":
            # Two framebuffers per rank (frame parity): peers may already store frame k+1 while this rank still reads
            # frame k; a rank re-enters a buffer only after one more completion all-reduce, i.e. after every reader's
            # stream has moved past that frame.
            lib, nbytes = tracer.lib, X * Y * 12
            self.frames, self.dst = [], []
            problem = None
            for parity in range(2):
                ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
                try:
                    A.check(lib.ft_device_alloc(tracer.device, nbytes, C.byref(ptr)))
                    self._owned.append(ptr)
                    A.check(lib.ft_ipc_export(tracer.device, ptr, handle))
                except A.FrayTracerError as e:
                    problem = problem or str(e)
                handles = [None] * self.world
                dist.all_gather_object(handles, (bytes(handle), problem), group=group)  # collective: every rank takes part
                problem = problem or next((p for _, p in handles if p), None)
                ptrs = []
                for r, (h, _) in enumerate(handles):
                    if problem:
                        break
                    if r == self.rank:
                        ptrs.append(ptr.value)
                        continue
                    peer = C.c_void_p()
                    try:
                        A.check(lib.ft_ipc_import(tracer.device, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(peer)))
                        self._imported.append(peer)
                        ptrs.append(peer.value)
                    except A.FrayTracerError as e:
                        problem = str(e)
                self.dst.append(ptrs)
                if not problem:
                    self.frames.append(torch.as_tensor(_DeviceArray(ptr.value, (X, Y, 3)), device=dev))
            # peer mapping must work on EVERY rank; otherwise all ranks use the NCCL all-gather (still a GPU path)
            votes = [None] * self.world
            dist.all_gather_object(votes, problem, group=group)
            problem = next((p for p in votes if p), None)
            if problem is None:
                self.final = self.frames[0]
                self._done = torch.zeros(1, dtype=torch.int32, device=dev)
                self._frame = 0
                self.launches_per_frame = 1
                return
            import warnings

            warnings.warn(f"CUDA IPC peer mapping unavailable ({problem}); gathering the framebuffer with NCCL instead")
            for p in self._imported:
                lib.ft_ipc_close(tracer.device, p)
            dist.barrier(group=group)
            for p in self._owned:
                lib.ft_device_free(tracer.device, p)
            self._owned, self._imported, self.frames, self.dst = [], [], [], []
            self.gather = "nccl"
        self.final = torch.empty((X, Y, 3), dtype=torch.float32, device=dev)
        if self.world > 1:
            chunks = max(1, min(int(chunks), self.spr))
            edges = [round(i * self.spr / chunks) for i in range(chunks + 1)]
            self.chunk_ranges = [(a, b - a) for a, b in zip(edges[:-1], edges[1:]) if b > a]
            self.local = [torch.empty((n * strip_width, Y, 3), dtype=torch.float32, device=dev)
                          for _, n in self.chunk_ranges]
            self.gathered = [torch.empty((self.world, n * strip_width, Y, 3), dtype=torch.float32, device=dev)
                             for _, n in self.chunk_ranges]
        self.launches_per_frame = 1 if self.world == 1 else 2 * len(self.chunk_ranges)

    # ---- the extension paths (configs 3 and 4: no reference implementation) over the same strips and peer framebuffers
    def render_extension(self, kind: str, camera, epsilon: float, length: float, **kw):
        """kind = "spectral" | "cloud": one frame on all ranks, every rank ends with the full framebuffer (P2P gather only).
        Returns (frame tensor, this rank's stats)."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            dst = [self.final.data_ptr()]
        else:
            if self.gather != "p2p":
                raise RuntimeError("the extension paths gather through peer framebuffers only (gather='p2p')")
            parity = self._frame & 1
            self._frame += 1
            dst = self.dst[parity]
        # strip_width 0: 4 x 8 pixel tiles dealt round-robin -- glass and cloud pixels cost tens of times a solid one, column
        # strips balance them badly (measured on 8 GPUs)
        st = self.tracer.render_extension_gather_device(kind, camera, self.size, epsilon, length, dst, stream=stream,
                                                        strip_width=kw.pop("strip_width", 0), rank=self.rank, world=self.world,
                                                        **kw)
        if self.world > 1:
            self.dist.all_reduce(self._done, group=self.group)
            self.final = self.frames[parity]
        return self.final, st

    # ---- framebuffer in shared, page-locked host memory: every GPU delivers its own strips over its own PCIe link
    def host_frame(self):
        """Creates (once) a FColor[X,Y] in POSIX shared memory, mapped and page-locked by every rank's process.
        Returns the numpy view; `render_to_host` fills it."""
        if getattr(self, "_host", None) is not None:
            return self._host
        import numpy as np
        from multiprocessing import shared_memory

        X, Y = self.size.X, self.size.Y
        nbytes = X * Y * 12
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
            _interleave_numa(self._shm.buf, nbytes)  # before the first touch (the page-locking below)
        if self.world > 1:
            self.dist.broadcast_object_list(name, src=0, group=self.group)
            if self.rank != 0:
                self._shm = shared_memory.SharedMemory(name=name[0])
                try:  # only the creating rank owns (and unlinks) the segment
                    from multiprocessing import resource_tracker

                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
        self._host = np.ndarray((X, Y, 3), dtype=np.float32, buffer=self._shm.buf)
        lib = self.tracer.lib
        self._host_addr = self._host.ctypes.data
        self.torch.cuda.set_device(self.tracer.device)
        A.check(lib.ft_host_register(C.c_void_p(self._host_addr), nbytes))
        dptr = C.c_void_p()
        A.check(lib.ft_host_get_device_pointer(C.c_void_p(self._host_addr), C.byref(dptr)))
        self._host_dev = dptr.value
        if not hasattr(self, "_done"):
            self._done = self.torch.zeros(1, dtype=self.torch.int32, device=self.torch.device("cuda", self.tracer.device))
        return self._host

    def render_to_host(self, camera, epsilon: float, length: float, flags: int = 0):
        """One frame straight into the shared host framebuffer (no device framebuffer, no device->host copy): every
        rank's kernel stores its strips over its own PCIe link; returns after all ranks have finished (the numpy view
        of `host_frame()` is then complete on every rank's host)."""
        host = self.host_frame()
        stream = self.torch.cuda.current_stream().cuda_stream
        self.tracer.render_gather_device(camera, self.size, epsilon, length, [self._host_dev], stream=stream, flags=flags,
                                         strip_width=self.strip_width, rank=self.rank, world=self.world)
        if self.world > 1:
            self.dist.all_reduce(self._done, group=self.group)
        self.torch.cuda.current_stream().synchronize()
        return host

    # ---- Image.toColors + Image.toBitmap over the ranks (Image.fs:37-72)
    def to_bitmap(self, frame, gamma: float, seed: int):
        """`frame`: the full FColor[X,Y] device tensor every rank holds after `render`.  Every rank takes the maximum over the
        columns of ITS strips (Array2D.maxWith, Image.fs:40-43), the maxima meet in an all-reduce(MAX), every rank tonemaps
        its own strips into a zeroed bitmap and -- each byte being written by exactly one rank -- an all-reduce(MAX) over
        the bitmaps assembles it.  Returns (uint8 [Y, X, 3] device tensor of top-down B,G,R scanlines in the reference's
        pixel order, identical on every rank; the raw maximum as a 1-element float32 device tensor)."""
        torch = self.torch
        X, Y = self.size.X, self.size.Y
        lib = self.tracer.lib
        stream = torch.cuda.current_stream().cuda_stream
        st = C.c_void_p(stream) if stream else None
        dev = frame.device
        mx = torch.empty(1, dtype=torch.float32, device=dev)
        bitmap = torch.zeros((Y, X, 3), dtype=torch.uint8, device=dev)
        A.check(lib.ft_image_max_strips_device(C.c_void_p(frame.data_ptr()), X, Y, self.strip_width, self.rank, self.world,
                                               C.c_void_p(mx.data_ptr()), st))
        if self.world > 1:
            self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX, group=self.group)
        A.check(lib.ft_tonemap_strips_device(C.c_void_p(frame.data_ptr()), X, Y, self.strip_width, self.rank, self.world,
                                             C.c_float(gamma), C.c_void_p(mx.data_ptr()), C.c_float(0.0), C.c_uint64(seed),
                                             C.c_void_p(bitmap.data_ptr()), 3 * X, st))
        if self.world > 1:
            self.dist.all_reduce(bitmap, op=self.dist.ReduceOp.MAX, group=self.group)
        return bitmap, mx

    def close(self):
        """release the IPC mappings and the shared framebuffers (after a barrier: no peer may still be writing)"""
        if getattr(self, "_host", None) is not None:
            self.torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier(group=self.group)
            self.tracer.lib.ft_host_unregister(C.c_void_p(self._host_addr))
            self._host = None
            self._shm.close()
            if self.rank == 0:
                self._shm.unlink()
        if self._owned:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            for p in self._imported:
                self.tracer.lib.ft_ipc_close(self.tracer.device, p)
            self.dist.barrier(group=self.group)
            for p in self._owned:
                self.tracer.lib.ft_device_free(self.tracer.device, p)
            self._owned, self._imported, self.frames = [], [], []

    def render(self, camera, epsilon: float, length: float, flags: int = 0, copy: bool = False):
        """One frame on all ranks; returns the full FColor[X,Y] as a device tensor on every rank.

        Lifetime of the returned tensor: with gather="p2p" it is a view of one of the rank's two IPC framebuffers (frame
        parity), which the peers overwrite two `render` calls later -- read it on the current stream before the second next
        call, or pass copy=True to get a private clone.  With gather="nccl" (and with one rank) it is `self.final`, owned by
        this renderer and overwritten by the next call."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        X, Y = self.size.X, self.size.Y
        if self.world > 1 and self.gather == "p2p":
            parity = self._frame & 1
            self._frame += 1
            self.tracer.render_gather_device(camera, self.size, epsilon, length, self.dst[parity], stream=stream, flags=flags,
                                             strip_width=self.strip_width, rank=self.rank, world=self.world)
            self.dist.all_reduce(self._done, group=self.group)  # every writer's kernel has finished: the frame is complete
            self.final = self.frames[parity]
            return self.final.clone() if copy else self.final
        if self.world == 1:
            self.tracer.render_device(camera, self.size, epsilon, length, self.final.data_ptr(), stream=stream,
                                      flags=flags, strip_width=self.strip_width, rank=0, world=1)
            return self.final
        works = []
        for (first, n), local, gathered in zip(self.chunk_ranges, self.local, self.gathered):
            self.tracer.render_device(camera, self.size, epsilon, length, local.data_ptr(), stream=stream, flags=flags,
                                      strip_width=self.strip_width, rank=self.rank, world=self.world,
                                      first_local_strip=first, n_local_strips=n)
            works.append(self.dist.all_gather_into_tensor(gathered.view(-1, Y, 3), local, group=self.group,
                                                          async_op=True))  # NCCL over NVLink, overlaps the next chunk
        for (first, n), gathered, work in zip(self.chunk_ranges, self.gathered, works):
            work.wait()  # the current stream waits for this chunk's collective
            A.check(self.tracer.lib.ft_deinterleave_strips_device(
                C.c_void_p(gathered.data_ptr()), X, Y, self.strip_width, self.world, 3, first, n,
                C.c_void_p(self.final.data_ptr()), C.c_void_p(stream) if stream else None))
        return self.final
