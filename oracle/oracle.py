"""ctypes wrapper of oracle/libft_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (fraytracer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fraytracer_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libft_oracle.so")


class fto_counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "march_steps", "primary_rays", "primary_hits", "shadow_rays",
        "eval_sphere", "eval_capsule", "eval_torus", "eval_triangle_edge", "eval_triangle_face",
        "union_lookups", "lower_bound_tests", "bound_tests", "intersect_tests", "max_steps",
        "outside_lookups", "outside_bound_tests", "outside_evals")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force: bool = False):
    src = os.path.join(HERE, "ft_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "libft_oracle.so"])
    return LIB_PATH


_libs = {}


def assumption_library(bits: int) -> str:
    """path of the oracle built with FTO_ASSUME=bits (ft_oracle.c: the unpinned assumptions about System.Numerics / Array.Sort);
    built on demand"""
    path = os.path.join(HERE, f"libft_oracle_assume{int(bits)}.so")
    src = os.path.join(HERE, "ft_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", os.path.basename(path)])
    return path


def lib(path: str | None = None):
    key = path or LIB_PATH
    if key not in _libs:
        if path is None and not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(key)
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        P = C.POINTER
        L.fto_random_create.restype = vp
        L.fto_random_create.argtypes = [i32]
        L.fto_random_destroy.argtypes = [vp]
        L.fto_random_next.restype = i32
        L.fto_random_next.argtypes = [vp]
        L.fto_random_next_double.restype = C.c_double
        L.fto_random_next_double.argtypes = [vp]
        L.fto_dotnet_sort.argtypes = [vp, vp, i32]
        L.fto_scene_create.restype = C.c_int
        L.fto_scene_create.argtypes = [P(A.ft_scene_desc), P(vp), C.c_char_p, C.c_int]
        L.fto_scene_destroy.argtypes = [vp]
        L.fto_scene_set_step_cap.argtypes = [vp, i64]
        L.fto_grid_count.restype = C.c_int
        L.fto_grid_count.argtypes = [vp]
        L.fto_grid_info.argtypes = [vp, C.c_int, P(A.ft_grid_info)]
        L.fto_grid_read.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.fto_lens_create.restype = f32
        L.fto_lens_create.argtypes = [f32]
        L.fto_camera_look_at.argtypes = [P(f32), P(f32), P(f32), f32, P(A.ft_camera)]
        L.fto_render.argtypes = [vp, P(A.ft_camera), C.c_int, C.c_int, f32, f32, C.c_int, C.c_int, C.c_int, C.c_int,
                                 vp, vp, vp, vp, P(fto_counters)]
        L.fto_trace_rays.argtypes = [vp, i64, vp, C.c_int, vp, vp, vp, vp, P(fto_counters)]
        L.fto_distance.argtypes = [vp, i64, vp, C.c_int, vp, P(fto_counters)]
        L.fto_form_distance.argtypes = [vp, C.c_int, i64, vp, vp]
        L.fto_form_boundary.argtypes = [vp, C.c_int, vp]
        L.fto_root_boundary.argtypes = [vp, vp]
        L.fto_philox4x32_10.argtypes = [vp, vp, vp]
        L.fto_render_cloud.argtypes = [vp, P(A.ft_camera), C.c_int, C.c_int, f32, f32, P(A.ft_cloud_params), C.c_int,
                                       C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.fto_noise_gradient.restype = C.c_double
        L.fto_noise_gradient.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double]
        L.fto_render_spectral.argtypes = [vp, P(A.ft_camera), C.c_int, C.c_int, f32, f32, C.c_int, C.c_int, C.c_uint64,
                                          C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.fto_to_bitmap.restype = f32
        L.fto_to_bitmap.argtypes = [vp, C.c_int, C.c_int, f32, C.c_uint64, vp, C.c_int]
        L.fto_assumptions.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleRandom:
    def __init__(self, seed):
        self.h = lib().fto_random_create(seed)

    def Next(self):
        return lib().fto_random_next(self.h)

    def NextDouble(self):
        return lib().fto_random_next_double(self.h)

    def __del__(self):
        lib().fto_random_destroy(self.h)


def dotnet_sort(keys, items):
    k = np.ascontiguousarray(keys, np.float32).copy()
    v = np.ascontiguousarray(items, np.int32).copy()
    lib().fto_dotnet_sort(_ptr(k), _ptr(v), len(k))
    return k, v


def lens_create(fov):
    return float(lib().fto_lens_create(C.c_float(fov)))


def camera_look_at(position, look_at, up, nps) -> A.ft_camera:
    out = A.ft_camera()
    arr = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    lib().fto_camera_look_at(arr(position), arr(look_at), arr(up), C.c_float(nps), C.byref(out))
    return out


class OracleScene:
    """CPU restatement of a scene, built from the same flat description the product takes."""

    def __init__(self, description, library: str | None = None):
        """description: fraytracer_b200.SceneDescription; library: another build of the oracle (assumption_library)"""
        self._L = lib(library)
        self.description = description
        self.h = C.c_void_p()
        err = C.create_string_buffer(256)
        st = self._L.fto_scene_create(C.byref(description.desc), C.byref(self.h), err, 256)
        if st != 0:
            raise A.FrayTracerError(st, err.value.decode())

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self._L.fto_scene_destroy(self.h)
            self.h = C.c_void_p()

    def set_step_cap(self, cap):
        self._L.fto_scene_set_step_cap(self.h, int(cap))

    def grids(self):
        res = []
        for g in range(self._L.fto_grid_count(self.h)):
            gi = A.ft_grid_info()
            self._L.fto_grid_info(self.h, g, C.byref(gi))
            nc = gi.count[0] * gi.count[1] * gi.count[2]
            centers = np.empty((nc, 3), np.float32)
            begin = np.empty(nc + 1, np.int64)
            lb = np.empty(gi.n_entries, np.float32)
            item = np.empty(gi.n_entries, np.int32)
            self._L.fto_grid_read(self.h, g, _ptr(centers), _ptr(begin), _ptr(lb), _ptr(item))
            res.append(dict(count=tuple(gi.count), aabb_min=np.array(gi.aabb_min, np.float32),
                            cell_size=np.array(gi.cell_size, np.float32),
                            cell_size_inv=np.array(gi.cell_size_inv, np.float32),
                            centers=centers, begin=begin, lower_bound=lb, item=item))
        return res

    def render(self, camera, X, Y, eps, length, threads=None, x_begin=0, x_end=None, x_stride=1, counters=False):
        """returns dict(rgb [X,Y,3], leaf [X,Y], steps [X,Y], shadow [X,Y]) (+ counters)"""
        cam = camera.to_c() if hasattr(camera, "to_c") else camera
        x_end = X if x_end is None else x_end
        threads = threads or os.cpu_count() or 1
        rgb = np.zeros((X, Y, 3), np.float32)
        leaf = np.full((X, Y), -2, np.int32)
        steps = np.zeros((X, Y), np.int32)
        shadow = np.zeros((X, Y), np.uint32)
        ctr = fto_counters()
        st = self._L.fto_render(self.h, C.byref(cam), X, Y, C.c_float(eps), C.c_float(length), x_begin, x_end,
                              x_stride, threads, _ptr(rgb), _ptr(leaf), _ptr(steps), _ptr(shadow),
                              C.byref(ctr) if counters else None)
        assert st == 0
        out = dict(rgb=rgb, leaf=leaf, steps=steps, shadow=shadow)
        if counters:
            out["counters"] = ctr.as_dict()
        return out

    def render_spectral(self, camera, X, Y, eps, length, n_wavelengths=16, max_bounces=8, seed=19, threads=None,
                        x_begin=0, x_end=None, x_stride=1):
        """EXTENSION (no reference implementation): CPU restatement of ft_render_spectral
        (only the columns x_begin, x_begin + x_stride, ... < x_end of the full-size array are written)"""
        cam = camera.to_c() if hasattr(camera, "to_c") else camera
        rgb = np.zeros((X, Y, 3), np.float32)
        st = self._L.fto_render_spectral(self.h, C.byref(cam), X, Y, C.c_float(eps), C.c_float(length), n_wavelengths,
                                       max_bounces, C.c_uint64(seed), x_begin, X if x_end is None else x_end, x_stride,
                                       threads or os.cpu_count() or 1, _ptr(rgb))
        assert st == 0, st
        return rgb

    def render_cloud(self, camera, X, Y, eps, length, cloud, spp=64, seed=19, threads=None, x_begin=0, x_end=None,
                     x_stride=1):
        """EXTENSION (no reference implementation): CPU restatement of ft_render_cloud.
        Returns (rgb [X,Y,3], dict(secondary, track_steps, scattered))"""
        cam = camera.to_c() if hasattr(camera, "to_c") else camera
        cp = cloud.to_c()
        rgb = np.zeros((X, Y, 3), np.float32)
        ctr = np.zeros(3, np.int64)
        st = self._L.fto_render_cloud(self.h, C.byref(cam), X, Y, C.c_float(eps), C.c_float(length), C.byref(cp), spp,
                                    C.c_uint64(seed), x_begin, X if x_end is None else x_end, x_stride,
                                    threads or os.cpu_count() or 1, _ptr(rgb), _ptr(ctr))
        assert st == 0, st
        return rgb, dict(secondary=int(ctr[0]), track_steps=int(ctr[1]), scattered=int(ctr[2]))

    def trace_rays(self, rays, threads=None, counters=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        threads = threads or os.cpu_count() or 1
        rgb = np.empty((n, 3), np.float32)
        leaf = np.empty(n, np.int32)
        steps = np.empty(n, np.int32)
        shadow = np.empty(n, np.uint32)
        ctr = fto_counters()
        st = self._L.fto_trace_rays(self.h, n, _ptr(rays), threads, _ptr(rgb), _ptr(leaf), _ptr(steps),
                                  _ptr(shadow), C.byref(ctr) if counters else None)
        assert st == 0
        out = dict(rgb=rgb, leaf=leaf, steps=steps, shadow=shadow)
        if counters:
            out["counters"] = ctr.as_dict()
        return out

    def distance(self, points, threads=None):
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.empty(points.shape[0], np.float32)
        st = self._L.fto_distance(self.h, points.shape[0], _ptr(points), threads or os.cpu_count() or 1,
                                _ptr(out), None)
        assert st == 0
        return out

    def form_distance(self, form_id, points):
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.empty(points.shape[0], np.float32)
        st = self._L.fto_form_distance(self.h, form_id, points.shape[0], _ptr(points), _ptr(out))
        assert st == 0, st
        return out

    def form_boundary(self, form_id):
        out = np.empty(4, np.float32)
        st = self._L.fto_form_boundary(self.h, form_id, _ptr(out))
        assert st == 0, st
        return out

    def root_boundary(self):
        out = np.empty(4, np.float32)
        self._L.fto_root_boundary(self.h, _ptr(out))
        return out


def noise_gradient(perm, gradients, x, y, z):
    """Noise.gradient x y z (Noise.fs:97-110) over the given tables"""
    perm = np.ascontiguousarray(perm, np.int32)
    gradients = np.ascontiguousarray(gradients, np.float64)
    return float(lib().fto_noise_gradient(_ptr(perm), _ptr(gradients), float(x), float(y), float(z)))


def philox4x32_10(ctr, key):
    c = np.array(ctr, np.uint32)
    k = np.array(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().fto_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def to_bitmap(rgb, gamma, seed):
    """Image.toColors + toBitmap pixel order: returns (uint8 [Y, X, 3] top-down B,G,R scanlines, raw max)"""
    rgb = np.ascontiguousarray(rgb, np.float32)
    X, Y = rgb.shape[0], rgb.shape[1]
    out = np.zeros((Y, X, 3), np.uint8)
    mx = lib().fto_to_bitmap(_ptr(rgb), X, Y, C.c_float(gamma), C.c_uint64(seed), _ptr(out), 3 * X)
    return out, float(mx)
