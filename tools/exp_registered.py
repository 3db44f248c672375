"""end to end through ft_render: cudaHostAlloc'd frame (ft_host_alloc) vs a caller's own array page-locked with ft_host_register
(what the F# FrameBuffer does) vs the same in POSIX shared memory (what DistributedRenderer.host_frame does)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fraytracer_b200 as ft
from fraytracer_b200 import scenes, _abi as A
from multiprocessing import shared_memory

for make in (scenes.spheres_capsules_scene, lambda: scenes.mixed_scene(width=3840, height=2160), scenes.mixed_scene):
    w = make(); X, Y = w.size.X, w.size.Y; nbytes = X * Y * 12
    tr = ft.SceneTracer(w.scene, 0); lib = tr.lib
    def run(out):
        ts = []
        for it in range(24):
            t0 = time.perf_counter(); tr.render(w.camera, w.size, w.epsilon, w.length, out=out); ts.append(time.perf_counter() - t0)
        ts = sorted(ts[6:]); return ts[len(ts) // 2] * 1e3
    host = C.c_void_p(); A.check(lib.ft_host_alloc(C.byref(host), nbytes))
    a = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(X, Y, 3)); t_alloc = run(a); A.check(lib.ft_host_free(host))
    raw = np.empty(nbytes + 4096, np.uint8); off = (-raw.ctypes.data) % 4096
    b = raw[off:off + nbytes].view(np.float32).reshape(X, Y, 3); b.fill(0)
    A.check(lib.ft_host_register(C.c_void_p(b.ctypes.data), nbytes)); t_reg = run(b); A.check(lib.ft_host_unregister(C.c_void_p(b.ctypes.data)))
    shm = shared_memory.SharedMemory(create=True, size=nbytes)
    c = np.ndarray((X, Y, 3), dtype=np.float32, buffer=shm.buf); c.fill(0)
    A.check(lib.ft_host_register(C.c_void_p(c.ctypes.data), nbytes)); t_shm = run(c); A.check(lib.ft_host_unregister(C.c_void_p(c.ctypes.data)))
    del c; shm.close(); shm.unlink()
    print(f"{w.name} {X}x{Y} ({nbytes/1e6:.0f} MB): cudaHostAlloc {t_alloc:.3f} ms | registered numpy array {t_reg:.3f} ms | registered POSIX shm {t_shm:.3f} ms")
    tr.close()
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| shmem:", open("/sys/kernel/mm/transparent_hugepage/shmem_enabled").read().strip())
