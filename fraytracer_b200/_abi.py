"""ctypes view of include/fraytracer_b200.h -- struct layouts, constants, loader.

The product path has no CPU fallback: `load_library()` raises when the in-tree
CUDA library is missing, and every compute entry point of the library itself
fails with FT_ERR_NO_DEVICE when no GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libfraytracer_b200.so")

FT_OK = 0
FT_ERR_INVALID_ARGUMENT = 1
FT_ERR_EMPTY_COMPOSITE = 2
FT_ERR_UNSUPPORTED_NODE = 3
FT_ERR_CUDA = 4
FT_ERR_DEPTH = 5
FT_ERR_NO_DEVICE = 6
FT_ERR_OUT_OF_MEMORY = 7

FT_DEVICE_NONE = -1
FT_MAX_DEPTH = 12
FT_MAX_LIGHTS = 32

FT_FORM_SPHERE = 1
FT_FORM_CAPSULE = 2
FT_FORM_TORUS = 3
FT_FORM_TRIANGLE = 4
FT_FORM_CONE = 5  # EXTENSION (no reference implementation): capped cone
FT_FORM_UNION = 16
FT_FORM_SUBTRACT = 17
FT_FORM_INTERSECT = 18
FT_FORM_UNION_SMOOTH = 19

FT_MATERIAL_SOLID = 1
FT_MATERIAL_GLASS = 2  # EXTENSION (no reference implementation): Cauchy glass, rgb = (A, B, unused)
FT_SPECTRAL_MAX_LIGHTS = 8
FT_MATERIAL_CLOUD = 3  # EXTENSION (no reference implementation): participating medium, rgb = albedo

FT_OBJECT_LEAF = 1
FT_OBJECT_UNION = 2
FT_OBJECT_SUBTRACT = 3
FT_OBJECT_INTERSECT = 4

FT_LIGHT_DIRECTIONAL = 1
FT_LIGHT_POINT = 2

FT_FLAG_DEFAULT = 0
FT_FLAG_COUNTERS = 1
FT_FLAG_FAST_MATH = 2


class ft_form_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("first_child", C.c_int32), ("child_count", C.c_int32),
                ("p", C.c_float * 10)]


class ft_material_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rgb", C.c_float * 3)]


class ft_object_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("form", C.c_int32), ("material", C.c_int32),
                ("first_child", C.c_int32), ("child_count", C.c_int32)]


class ft_light_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("v", C.c_float * 3), ("rgb", C.c_float * 3)]


class ft_scene_desc(C.Structure):
    _fields_ = [("forms", C.POINTER(ft_form_desc)), ("objects", C.POINTER(ft_object_desc)),
                ("children", C.POINTER(C.c_int32)), ("materials", C.POINTER(ft_material_desc)),
                ("lights", C.POINTER(ft_light_desc)),
                ("n_forms", C.c_int32), ("n_objects", C.c_int32), ("n_children", C.c_int32),
                ("n_materials", C.c_int32), ("n_lights", C.c_int32), ("root_object", C.c_int32),
                ("background", C.c_float * 3)]


class ft_camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("forward", C.c_float * 3),
                ("up_scaled", C.c_float * 3), ("right_scaled", C.c_float * 3)]


class ft_ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3),
                ("length", C.c_float), ("epsilon", C.c_float)]


class ft_scene_info(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_primitives", C.c_int32), ("n_grids", C.c_int32),
                ("max_depth", C.c_int32), ("n_cells", C.c_int64), ("n_entries", C.c_int64),
                ("device_bytes", C.c_int64), ("device", C.c_int32), ("sm_count", C.c_int32),
                ("grids_built_on_gpu", C.c_int32), ("reserved", C.c_int32)]


class ft_grid_info(C.Structure):
    _fields_ = [("node", C.c_int32), ("count", C.c_int32 * 3), ("aabb_min", C.c_float * 3),
                ("cell_size", C.c_float * 3), ("cell_size_inv", C.c_float * 3),
                ("n_entries", C.c_int64)]


class ft_subgrid_info(C.Structure):
    _fields_ = [("resolution", C.c_int32), ("halo", C.c_int32), ("n_subcells", C.c_int64), ("n_entries", C.c_int64),
                ("device_bytes", C.c_int64), ("threshold", C.c_float), ("build_ms", C.c_float)]


class ft_cloud_params(C.Structure):
    _fields_ = [("density", C.c_float), ("frequency", C.c_float), ("step", C.c_float), ("fade", C.c_float),
                ("bias", C.c_float), ("gain", C.c_float), ("perm", C.POINTER(C.c_int32)),
                ("gradients", C.POINTER(C.c_double))]


class ft_render_stats(C.Structure):
    _fields_ = [("primary_rays", C.c_int64), ("primary_hits", C.c_int64), ("shadow_rays", C.c_int64),
                ("capped_rays", C.c_int64), ("march_steps", C.c_int64), ("eval_flops", C.c_int64),
                ("escaped_rays", C.c_int64),
                ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32),
                ("reserved", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/fraytracer_b200.h declares (tests check the library exports them all)
EXPORTS = [
    "ft_abi_version", "ft_last_error", "ft_device_count", "ft_lens_create", "ft_camera_look_at",
    "ft_scene_create", "ft_scene_destroy", "ft_scene_info_get", "ft_scene_set_step_cap",
    "ft_scene_grid_count", "ft_scene_grid_info", "ft_scene_grid_read", "ft_scene_subgrid_info", "ft_scene_set_debug",
    "ft_render", "ft_render_device", "ft_render_strips_device", "ft_render_gather_device",
    "ft_device_alloc", "ft_device_free", "ft_ipc_export", "ft_ipc_import", "ft_ipc_close", "ft_strip_layout", "ft_deinterleave_device",
    "ft_deinterleave_strips_device",
    "ft_trace_rays", "ft_distance", "ft_render_spectral", "ft_render_cloud", "ft_render_spectral_gather_device", "ft_render_cloud_gather_device",
    "ft_image_max_device", "ft_tonemap_device", "ft_image_max_strips_device", "ft_tonemap_strips_device", "ft_to_bitmap", "ft_save_bitmap",
    "ft_host_register", "ft_host_unregister", "ft_host_alloc", "ft_host_free", "ft_host_get_device_pointer",
    "ft_measure_fp32_peak",
]

_lib = None


class FrayTracerError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[fraytracer_b200 status {code}] {message}")
        self.code = code


def load_library(path: str | None = None):
    """Load libfraytracer_b200.so (in-tree build).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FT_LIB") or LIB_PATH  # FT_LIB: development experiments only
    if not os.path.exists(p):
        raise FrayTracerError(
            FT_ERR_NO_DEVICE,
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fraytracer_b200/csrc`). fraytracer_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    P = C.POINTER
    sig = {
        "ft_abi_version": (C.c_int, []),
        "ft_last_error": (C.c_char_p, []),
        "ft_device_count": (C.c_int, [P(C.c_int)]),
        "ft_lens_create": (f32, [f32]),
        "ft_camera_look_at": (C.c_int, [P(f32), P(f32), P(f32), f32, P(ft_camera)]),
        "ft_scene_create": (C.c_int, [P(ft_scene_desc), C.c_int, P(vp)]),
        "ft_scene_destroy": (C.c_int, [vp]),
        "ft_scene_info_get": (C.c_int, [vp, P(ft_scene_info)]),
        "ft_scene_set_step_cap": (C.c_int, [vp, i32]),
        "ft_scene_subgrid_info": (C.c_int, [vp, P(ft_subgrid_info)]),
        "ft_scene_set_debug": (C.c_int, [vp, i32]),
        "ft_scene_grid_count": (C.c_int, [vp, P(i32)]),
        "ft_scene_grid_info": (C.c_int, [vp, i32, P(ft_grid_info)]),
        "ft_scene_grid_read": (C.c_int, [vp, i32, vp, vp, vp, vp]),
        "ft_render": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, u32, vp, vp, P(ft_render_stats)]),
        "ft_render_device": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, u32, i32, i32, i32, vp, vp, vp,
                                       P(ft_render_stats)]),
        "ft_render_strips_device": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, u32, i32, i32, i32, i32, i32, vp, vp,
                                              vp, P(ft_render_stats)]),
        "ft_render_gather_device": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, u32, i32, i32, i32, P(vp), i32, vp,
                                              P(ft_render_stats)]),
        "ft_device_alloc": (C.c_int, [C.c_int, i64, P(vp)]),
        "ft_device_free": (C.c_int, [C.c_int, vp]),
        "ft_ipc_export": (C.c_int, [C.c_int, vp, vp]),
        "ft_ipc_import": (C.c_int, [C.c_int, vp, P(vp)]),
        "ft_ipc_close": (C.c_int, [C.c_int, vp]),
        "ft_deinterleave_strips_device": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
        "ft_strip_layout": (C.c_int, [i32, i32, i32, i32, P(i32), P(i64)]),
        "ft_deinterleave_device": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "ft_trace_rays": (C.c_int, [vp, i64, vp, u32, vp, vp, vp, vp, P(ft_render_stats)]),
        "ft_distance": (C.c_int, [vp, i64, vp, u32, vp]),
        "ft_render_spectral": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, i32, i32, C.c_uint64, vp,
                                         P(ft_render_stats)]),
        "ft_render_cloud": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, P(ft_cloud_params), i32, C.c_uint64, vp,
                                      P(ft_render_stats)]),
        "ft_render_spectral_gather_device": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, i32, i32, C.c_uint64, i32, i32, i32,
                                                       P(vp), i32, vp, P(ft_render_stats)]),
        "ft_render_cloud_gather_device": (C.c_int, [vp, P(ft_camera), i32, i32, f32, f32, P(ft_cloud_params), i32, C.c_uint64,
                                                    i32, i32, i32, P(vp), i32, vp, P(ft_render_stats)]),
        "ft_image_max_device": (C.c_int, [vp, i64, vp, vp]),
        "ft_tonemap_device": (C.c_int, [vp, i32, i32, f32, vp, f32, C.c_uint64, vp, i32, vp]),
        "ft_image_max_strips_device": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "ft_tonemap_strips_device": (C.c_int, [vp, i32, i32, i32, i32, i32, f32, vp, f32, C.c_uint64, vp, i32, vp]),
        "ft_to_bitmap": (C.c_int, [C.c_int, vp, i32, i32, f32, C.c_uint64, vp, i32]),
        "ft_save_bitmap": (C.c_int, [C.c_char_p, vp, i32, i32, i32]),
        "ft_host_register": (C.c_int, [vp, i64]),
        "ft_host_unregister": (C.c_int, [vp]),
        "ft_host_alloc": (C.c_int, [P(vp), i64]),
        "ft_host_free": (C.c_int, [vp]),
        "ft_host_get_device_pointer": (C.c_int, [vp, P(vp)]),
        "ft_measure_fp32_peak": (C.c_int, [C.c_int, P(f32), P(f32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(status: int):
    if status != FT_OK:
        lib = load_library()
        msg = lib.ft_last_error()
        raise FrayTracerError(status, msg.decode("utf-8", "replace") if msg else "unknown error")
