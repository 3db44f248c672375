"""Host-side mirror of the reference's extension API for the per-ray hot path.

Same constructor names, argument order and error behaviour as the F# modules
(`SdfForm.*`, `SdfObject.*`, `SdfMaterial.*`, `SdfLight.*`, `SdfScene.trace`,
`Camera.lookAt`, `Lens.create`, `Image.render`), but the constructors return
*reified* nodes instead of closures (closures cannot be compiled to a device
node buffer -- SURVEY.md section 8(b)).  `SdfScene.trace` hands the flat
description to the C-ABI scene compiler (`ft_scene_create`) and returns a
`SceneTracer`; `Image.render` is one `ft_render` call.

Reference: /root/reference/src/FrayTracer/{Types,SdfForm,SdfObject,SdfMaterial,
SdfLight,SdfScene,Camera,Image}.fs
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

from . import _abi as A

f32 = np.float32


def _vec3(v) -> tuple:
    a = np.asarray(v, dtype=np.float32).reshape(-1)
    if a.shape[0] != 3:
        raise ValueError("expected a 3-vector")
    return (float(a[0]), float(a[1]), float(a[2]))


# --------------------------------------------------------------------------- FColor
class FColor:
    """FColor.fs:7-34 -- RGB float32 triple."""

    @staticmethod
    def ofRGB(r, g, b) -> tuple:
        return (float(f32(r)), float(f32(g)), float(f32(b)))


# --------------------------------------------------------------------------- Ray
@dataclass(frozen=True)
class Ray:
    """Types.fs:9-17"""
    Origin: tuple
    Direction: tuple
    Length: float
    Epsilon: float

    def as_row(self):
        return np.array([*self.Origin, *self.Direction, self.Length, self.Epsilon], dtype=np.float32)


# --------------------------------------------------------------------------- SdfForm
class SdfForm:
    """Reified SdfForm (Types.fs:40-44).  `kind` is an FT_FORM_* tag, `params` the
    constructor's record fields, `children` the child forms of a composite."""

    __slots__ = ("kind", "params", "children")

    def __init__(self, kind: int, params: Sequence[float] = (), children: Sequence["SdfForm"] = ()):
        self.kind = kind
        self.params = tuple(float(f32(x)) for x in params)
        self.children = tuple(children)

    # SdfForm.fs:14-40
    @staticmethod
    def union(forms: Iterable["SdfForm"]) -> "SdfForm":
        forms = list(forms)
        if not forms:
            raise ValueError("No SdfObjects given.")  # SdfForm.fs:16
        if len(forms) == 1:
            return forms[0]  # SdfForm.fs:17
        return SdfForm(A.FT_FORM_UNION, (), forms)

    # SdfForm.fs:42-49
    @staticmethod
    def subtract(a: "SdfForm", b: "SdfForm") -> "SdfForm":
        return SdfForm(A.FT_FORM_SUBTRACT, (), (a, b))

    # SdfForm.fs:51-67
    @staticmethod
    def intersect(forms: Iterable["SdfForm"]) -> "SdfForm":
        forms = list(forms)
        if not forms:
            raise ValueError("No SdfObjects given.")  # SdfForm.fs:53
        if len(forms) == 1:
            return forms[0]
        return SdfForm(A.FT_FORM_INTERSECT, (), forms)

    # SdfForm.fs:69-91
    @staticmethod
    def unionSmooth(strength: float, forms: Iterable["SdfForm"]) -> "SdfForm":
        forms = list(forms)
        if not forms:
            raise ValueError("blub")  # SdfForm.fs:71
        if len(forms) == 1:
            return forms[0]
        return SdfForm(A.FT_FORM_UNION_SMOOTH, (strength,), forms)

    class Primitive:
        """SdfForm.fs:117-268"""

        @dataclass(frozen=True)
        class Sphere:
            Center: tuple
            Radius: float

        @dataclass(frozen=True)
        class Capsule:
            From: tuple
            To: tuple
            Radius: float

        @dataclass(frozen=True)
        class Torus:
            Center: tuple
            Normal: tuple
            MajorRadius: float
            MinorRadius: float

        @dataclass(frozen=True)
        class Triangle:
            V1: tuple
            V2: tuple
            V3: tuple
            Radius: float

        @dataclass(frozen=True)
        class Cone:
            """EXTENSION, not in the reference (BASELINE.json names cones, the reference has none): capped cone"""
            From: tuple
            To: tuple
            FromRadius: float
            ToRadius: float

        @staticmethod
        def cone(data=None, **kw) -> "SdfForm":
            d = data or SdfForm.Primitive.Cone(**kw)
            return SdfForm(A.FT_FORM_CONE, (*_vec3(d.From), *_vec3(d.To), d.FromRadius, d.ToRadius))

        @staticmethod
        def sphere(data=None, **kw) -> "SdfForm":
            d = data or SdfForm.Primitive.Sphere(**kw)
            return SdfForm(A.FT_FORM_SPHERE, (*_vec3(d.Center), d.Radius))

        @staticmethod
        def capsule(data=None, **kw) -> "SdfForm":
            d = data or SdfForm.Primitive.Capsule(**kw)
            return SdfForm(A.FT_FORM_CAPSULE, (*_vec3(d.From), *_vec3(d.To), d.Radius))

        @staticmethod
        def torus(data=None, **kw) -> "SdfForm":
            d = data or SdfForm.Primitive.Torus(**kw)
            return SdfForm(A.FT_FORM_TORUS, (*_vec3(d.Center), *_vec3(d.Normal), d.MajorRadius, d.MinorRadius))

        @staticmethod
        def triangle(data=None, **kw) -> "SdfForm":
            d = data or SdfForm.Primitive.Triangle(**kw)
            return SdfForm(A.FT_FORM_TRIANGLE, (*_vec3(d.V1), *_vec3(d.V2), *_vec3(d.V3), d.Radius))


# --------------------------------------------------------------------------- SdfMaterial
class SdfMaterial:
    """Types.fs:46-49.  Only constructor in the reference: createSolid (SdfMaterial.fs:4-7)."""

    __slots__ = ("kind", "rgb")

    def __init__(self, kind, rgb):
        self.kind = kind
        self.rgb = tuple(float(f32(x)) for x in rgb)

    @staticmethod
    def createSolid(color) -> "SdfMaterial":
        return SdfMaterial(A.FT_MATERIAL_SOLID, _vec3(color))

    @staticmethod
    def createCloud(albedo) -> "SdfMaterial":
        """EXTENSION, not in the reference: participating medium bounded by the object's form, `albedo`
        its single-scattering colour.  Only `SceneTracer.render_cloud` renders it."""
        return SdfMaterial(A.FT_MATERIAL_CLOUD, _vec3(albedo))

    @staticmethod
    def createGlass(cauchy_a: float, cauchy_b: float) -> "SdfMaterial":
        """EXTENSION, not in the reference: dispersive glass with the Cauchy index
        n(lambda) = A + B / lambda^2 (lambda in micrometres).  Only `SceneTracer.render_spectral` renders it."""
        return SdfMaterial(A.FT_MATERIAL_GLASS, (cauchy_a, cauchy_b, 0.0))


# --------------------------------------------------------------------------- SdfObject
class SdfObject:
    """Reified SdfObject (Types.fs:51-55, SdfObject.fs)."""

    __slots__ = ("kind", "form", "material", "base", "children", "forms")

    def __init__(self, kind, form=None, material=None, base=None, children=(), forms=()):
        self.kind = kind
        self.form = form          # LEAF: its form; SUBTRACT: the subtracted form
        self.material = material  # LEAF
        self.base = base          # SUBTRACT / INTERSECT
        self.children = tuple(children)  # UNION
        self.forms = tuple(forms)        # INTERSECT: the extra forms

    # SdfObject.fs:6-10 (argument order: material, then form)
    @staticmethod
    def create(material: SdfMaterial, form: SdfForm) -> "SdfObject":
        if not isinstance(material, SdfMaterial) or not isinstance(form, SdfForm):
            raise TypeError("SdfObject.create material form: closures are not supported "
                            "(FT_ERR_UNSUPPORTED_NODE); use the reified constructors")
        return SdfObject(A.FT_OBJECT_LEAF, form=form, material=material)

    # SdfObject.fs:12-48
    @staticmethod
    def union(objects: Iterable["SdfObject"]) -> "SdfObject":
        objects = list(objects)
        if not objects:
            raise ValueError("No SdfObjects given.")  # SdfObject.fs:14
        if len(objects) == 1:
            return objects[0]  # SdfObject.fs:16
        return SdfObject(A.FT_OBJECT_UNION, children=objects)

    # SdfObject.fs:50-54
    @staticmethod
    def subtract(object: "SdfObject", form: SdfForm) -> "SdfObject":
        return SdfObject(A.FT_OBJECT_SUBTRACT, form=form, base=object)

    # SdfObject.fs:56-64
    @staticmethod
    def intersect(object: "SdfObject", forms: Iterable[SdfForm]) -> "SdfObject":
        return SdfObject(A.FT_OBJECT_INTERSECT, base=object, forms=list(forms))

    @property
    def Form(self) -> SdfForm:
        """The object's Form as the reference composes it (SdfObject.fs:18-21,52,58-62)."""
        if self.kind == A.FT_OBJECT_LEAF:
            return self.form
        if self.kind == A.FT_OBJECT_UNION:
            return SdfForm.union([c.Form for c in self.children])
        if self.kind == A.FT_OBJECT_SUBTRACT:
            return SdfForm.subtract(self.base.Form, self.form)
        return SdfForm.intersect([self.base.Form, *self.forms])


# --------------------------------------------------------------------------- SdfLight
class SdfLight:
    """SdfLight.fs:6-42"""

    __slots__ = ("kind", "v", "rgb")

    def __init__(self, kind, v, rgb):
        self.kind = kind
        self.v = _vec3(v)
        self.rgb = _vec3(rgb)

    @staticmethod
    def directional(direction, color) -> "SdfLight":
        return SdfLight(A.FT_LIGHT_DIRECTIONAL, direction, color)

    @staticmethod
    def point(position, color) -> "SdfLight":
        return SdfLight(A.FT_LIGHT_POINT, position, color)


# --------------------------------------------------------------------------- Camera
@dataclass(frozen=True)
class Lens:
    """Camera.fs:5-14"""
    NearPlaneSize: float

    @staticmethod
    def create(fieldOfView: float) -> "Lens":
        return Lens(float(A.load_library().ft_lens_create(C.c_float(fieldOfView))))


@dataclass(frozen=True)
class Camera:
    """Camera.fs:16-54"""
    Position: tuple
    Forward: tuple
    UpScaled: tuple
    RightScaled: tuple

    @staticmethod
    def lookAt(Position, LookAt, Up, Lens: Lens) -> "Camera":
        lib = A.load_library()
        out = A.ft_camera()
        arr = lambda v: (C.c_float * 3)(*_vec3(v))
        A.check(lib.ft_camera_look_at(arr(Position), arr(LookAt), arr(Up), C.c_float(Lens.NearPlaneSize),
                                      C.byref(out)))
        return Camera(tuple(out.position), tuple(out.forward), tuple(out.up_scaled), tuple(out.right_scaled))

    def to_c(self) -> A.ft_camera:
        c = A.ft_camera()
        c.position[:] = self.Position
        c.forward[:] = self.Forward
        c.up_scaled[:] = self.UpScaled
        c.right_scaled[:] = self.RightScaled
        return c


@dataclass(frozen=True)
class ImageSize:
    """Image.fs:9-14"""
    X: int
    Y: int


# --------------------------------------------------------------------------- SdfScene
class SceneDescription:
    """The flat, blittable ft_scene_desc of a scene plus the arrays that keep it alive."""

    def __init__(self, scene: "SdfScene"):
        forms, objects, children, materials = [], [], [], []
        form_ids, object_ids, material_ids = {}, {}, {}

        def add_form(f: SdfForm, depth=0) -> int:
            if not isinstance(f, SdfForm):
                raise TypeError("only reified SdfForm nodes can cross the boundary (FT_ERR_UNSUPPORTED_NODE)")
            if id(f) in form_ids:
                return form_ids[id(f)]
            if depth > 256:
                raise ValueError("form nesting too deep")
            kid_ids = [add_form(k, depth + 1) for k in f.children]
            d = A.ft_form_desc()
            d.kind = f.kind
            d.first_child = len(children)
            d.child_count = len(kid_ids)
            children.extend(kid_ids)
            for i, x in enumerate(f.params):
                d.p[i] = x
            forms.append(d)
            form_ids[id(f)] = len(forms) - 1
            return len(forms) - 1

        def add_material(m: SdfMaterial) -> int:
            if not isinstance(m, SdfMaterial):
                raise TypeError("only reified SdfMaterial values can cross the boundary (FT_ERR_UNSUPPORTED_NODE)")
            if id(m) in material_ids:
                return material_ids[id(m)]
            d = A.ft_material_desc()
            d.kind = m.kind
            d.rgb[:] = m.rgb
            materials.append(d)
            material_ids[id(m)] = len(materials) - 1
            return len(materials) - 1

        def add_object(o: SdfObject, depth=0) -> int:
            if not isinstance(o, SdfObject):
                raise TypeError("only reified SdfObject nodes can cross the boundary (FT_ERR_UNSUPPORTED_NODE)")
            if id(o) in object_ids:
                return object_ids[id(o)]
            if depth > 256:
                raise ValueError("object nesting too deep")
            d = A.ft_object_desc()
            d.kind = o.kind
            d.form = d.material = -1
            if o.kind == A.FT_OBJECT_LEAF:
                d.form = add_form(o.form)
                d.material = add_material(o.material)
                d.first_child, d.child_count = 0, 0
            elif o.kind == A.FT_OBJECT_UNION:
                ids = [add_object(c, depth + 1) for c in o.children]
                d.first_child, d.child_count = len(children), len(ids)
                children.extend(ids)
            elif o.kind == A.FT_OBJECT_SUBTRACT:
                ids = [add_object(o.base, depth + 1)]
                d.form = add_form(o.form)
                d.first_child, d.child_count = len(children), 1
                children.extend(ids)
            elif o.kind == A.FT_OBJECT_INTERSECT:
                ids = [add_object(o.base, depth + 1)] + [add_form(f) for f in o.forms]
                d.first_child, d.child_count = len(children), len(ids)
                children.extend(ids)
            else:
                raise ValueError("unknown object kind")
            objects.append(d)
            object_ids[id(o)] = len(objects) - 1
            return len(objects) - 1

        root = add_object(scene.Object)
        lights = []
        for l in scene.Lights:
            d = A.ft_light_desc()
            d.kind = l.kind
            d.v[:] = l.v
            d.rgb[:] = l.rgb
            lights.append(d)

        self.object_ids = object_ids
        self._forms = (A.ft_form_desc * max(1, len(forms)))(*forms)
        self._objects = (A.ft_object_desc * max(1, len(objects)))(*objects)
        self._children = (C.c_int32 * max(1, len(children)))(*children)
        self._materials = (A.ft_material_desc * max(1, len(materials)))(*materials)
        self._lights = (A.ft_light_desc * max(1, len(lights)))(*lights)
        d = A.ft_scene_desc()
        d.forms = C.cast(self._forms, C.POINTER(A.ft_form_desc))
        d.objects = C.cast(self._objects, C.POINTER(A.ft_object_desc))
        d.children = C.cast(self._children, C.POINTER(C.c_int32))
        d.materials = C.cast(self._materials, C.POINTER(A.ft_material_desc))
        d.lights = C.cast(self._lights, C.POINTER(A.ft_light_desc))
        d.n_forms, d.n_objects, d.n_children = len(forms), len(objects), len(children)
        d.n_materials, d.n_lights, d.root_object = len(materials), len(lights), root
        d.background[:] = _vec3(scene.BackgroundColor)
        self.desc = d

    def leaf_id(self, obj: SdfObject) -> int:
        """index reported in out_leaf for this leaf SdfObject"""
        return self.object_ids[id(obj)]


@dataclass
class SdfScene:
    """Types.fs:74-79"""
    Object: SdfObject
    BackgroundColor: tuple
    Lights: Sequence[SdfLight]

    @staticmethod
    def trace(scene: "SdfScene", device: int = 0) -> "SceneTracer":
        """SdfScene.trace scene (SdfScene.fs:7-28): returns the ray-test, here a compiled device scene."""
        return SceneTracer(scene, device)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SceneTracer:
    """A scene compiled onto one B200 (`ft_scene` handle).  Callable like the reference's
    `Ray -> FColor` closure; batched entry points are what the hot path uses."""

    def __init__(self, scene: SdfScene, device: int = 0):
        self.lib = A.load_library()
        self.description = SceneDescription(scene)
        self.handle = C.c_void_p()
        A.check(self.lib.ft_scene_create(C.byref(self.description.desc), int(device), C.byref(self.handle)))
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.ft_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- introspection
    def info(self) -> dict:
        out = A.ft_scene_info()
        A.check(self.lib.ft_scene_info_get(self.handle, C.byref(out)))
        return {n: getattr(out, n) for n, _ in out._fields_}

    def subgrid_info(self) -> dict:
        """the nearest-candidate sub-grid built over the chain union's spatial lookup (resolution 0: none)"""
        out = A.ft_subgrid_info()
        A.check(self.lib.ft_scene_subgrid_info(self.handle, C.byref(out)))
        return {n: getattr(out, n) for n, _ in out._fields_ if n != "reserved"}

    def set_debug(self, sub_debug: int):
        """test hook: bit 0 = distrust every sub-grid certificate (every evaluation also runs the reference's scan)"""
        A.check(self.lib.ft_scene_set_debug(self.handle, int(sub_debug)))

    def set_step_cap(self, max_steps: int):
        A.check(self.lib.ft_scene_set_step_cap(self.handle, int(max_steps)))

    def grids(self) -> list:
        n = C.c_int32()
        A.check(self.lib.ft_scene_grid_count(self.handle, C.byref(n)))
        res = []
        for g in range(n.value):
            gi = A.ft_grid_info()
            A.check(self.lib.ft_scene_grid_info(self.handle, g, C.byref(gi)))
            nc = gi.count[0] * gi.count[1] * gi.count[2]
            centers = np.empty((nc, 3), np.float32)
            begin = np.empty(nc + 1, np.int64)
            lb = np.empty(gi.n_entries, np.float32)
            item = np.empty(gi.n_entries, np.int32)
            A.check(self.lib.ft_scene_grid_read(self.handle, g, _ptr(centers), _ptr(begin), _ptr(lb), _ptr(item)))
            res.append(dict(node=gi.node, count=tuple(gi.count), aabb_min=np.array(gi.aabb_min, np.float32),
                            cell_size=np.array(gi.cell_size, np.float32),
                            cell_size_inv=np.array(gi.cell_size_inv, np.float32),
                            centers=centers, begin=begin, lower_bound=lb, item=item))
        return res

    # -- the hot path
    def render(self, camera: Camera, size: ImageSize, epsilon: float, length: float, flags: int = 0,
               want_leaf: bool = False, out: np.ndarray | None = None, out_leaf: np.ndarray | None = None):
        """Image.render: returns (FColor[X,Y] as float32 [X,Y,3], leaf [X,Y] or None, stats dict)."""
        X, Y = int(size.X), int(size.Y)
        if out is None:
            out = np.empty((X, Y, 3), np.float32)
        if want_leaf and out_leaf is None:
            out_leaf = np.empty((X, Y), np.int32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == X * Y * 3
        stats = A.ft_render_stats()
        cam = camera.to_c()
        A.check(self.lib.ft_render(self.handle, C.byref(cam), X, Y, C.c_float(epsilon), C.c_float(length),
                                   int(flags), _ptr(out), _ptr(out_leaf), C.byref(stats)))
        return out, out_leaf, stats.as_dict()

    def render_spectral(self, camera: Camera, size: ImageSize, epsilon: float, length: float, n_wavelengths: int = 16,
                        max_bounces: int = 8, seed: int = 19, out: np.ndarray | None = None):
        """EXTENSION (BASELINE config 3; no reference implementation): spectral ray-test with
        dispersive glass, `n_wavelengths` frequencies per pixel.  Returns (float32 [X,Y,3], stats dict);
        stats["shadow_rays"] counts every secondary ray (shadow rays + continuation segments), stats["reserved"]
        the pixels whose primary ray hit glass."""
        X, Y = int(size.X), int(size.Y)
        if out is None:
            out = np.empty((X, Y, 3), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == X * Y * 3
        stats = A.ft_render_stats()
        cam = camera.to_c()
        A.check(self.lib.ft_render_spectral(self.handle, C.byref(cam), X, Y, C.c_float(epsilon), C.c_float(length),
                                            int(n_wavelengths), int(max_bounces), C.c_uint64(seed), _ptr(out),
                                            C.byref(stats)))
        return out, stats.as_dict()

    def render_cloud(self, camera: Camera, size: ImageSize, epsilon: float, length: float, cloud: "CloudParams",
                     samples_per_pixel: int = 64, seed: int = 19, out: np.ndarray | None = None):
        """EXTENSION (BASELINE config 4; no reference implementation): stochastic volumetric render,
        `samples_per_pixel` jittered samples per pixel.  Returns (float32 [X,Y,3] mean image, stats dict):
        stats["shadow_rays"] counts every secondary march (shadow rays, continuation segments, light tracks),
        stats["march_steps"] the tracking steps, stats["reserved"] the samples that scattered inside a cloud."""
        X, Y = int(size.X), int(size.Y)
        if out is None:
            out = np.empty((X, Y, 3), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == X * Y * 3
        stats = A.ft_render_stats()
        cam = camera.to_c()
        cp = cloud.to_c()
        A.check(self.lib.ft_render_cloud(self.handle, C.byref(cam), X, Y, C.c_float(epsilon), C.c_float(length),
                                         C.byref(cp), int(samples_per_pixel), C.c_uint64(seed), _ptr(out),
                                         C.byref(stats)))
        return out, stats.as_dict()

    def render_device(self, camera: Camera, size: ImageSize, epsilon: float, length: float, d_out_rgb: int,
                      d_out_leaf: int = 0, stream: int = 0, flags: int = 0, strip_width: int = 8,
                      rank: int = 0, world: int = 1, want_stats: bool = False, first_local_strip: int = 0,
                      n_local_strips: int = -1):
        """Asynchronous render into device memory (raw pointers, e.g. torch `data_ptr()`).
        want_stats=True synchronises the stream and returns the counters of this launch.
        first_local_strip / n_local_strips render only a range of the rank's strips (d_out_rgb then points at
        that range's first column)."""
        cam = camera.to_c()
        stats = A.ft_render_stats() if want_stats else None
        A.check(self.lib.ft_render_strips_device(self.handle, C.byref(cam), int(size.X), int(size.Y),
                                                 C.c_float(epsilon), C.c_float(length), int(flags), int(strip_width),
                                                 int(rank), int(world), int(first_local_strip), int(n_local_strips),
                                                 C.c_void_p(d_out_rgb), C.c_void_p(d_out_leaf) if d_out_leaf else None,
                                                 C.c_void_p(stream) if stream else None,
                                                 C.byref(stats) if want_stats else None))
        return stats.as_dict() if want_stats else None

    def render_gather_device(self, camera: Camera, size: ImageSize, epsilon: float, length: float, dst_ptrs, stream: int = 0,
                             flags: int = 0, strip_width: int = 8, rank: int = 0, world: int = 1, want_stats: bool = False):
        """Fused render + framebuffer gather: renders the strips `rank` owns and stores every pixel at its final position
        in each framebuffer of `dst_ptrs` (raw device pointers: own buffer and IPC-mapped peers)."""
        cam = camera.to_c()
        stats = A.ft_render_stats() if want_stats else None
        arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
        A.check(self.lib.ft_render_gather_device(self.handle, C.byref(cam), int(size.X), int(size.Y), C.c_float(epsilon),
                                                 C.c_float(length), int(flags), int(strip_width), int(rank), int(world),
                                                 arr, len(dst_ptrs), C.c_void_p(stream) if stream else None,
                                                 C.byref(stats) if want_stats else None))
        return stats.as_dict() if want_stats else None

    def render_extension_gather_device(self, kind: str, camera: Camera, size: ImageSize, epsilon: float, length: float, dst_ptrs,
                                       stream: int = 0, strip_width: int = 8, rank: int = 0, world: int = 1, seed: int = 19,
                                       n_wavelengths: int = 16, max_bounces: int = 8, cloud: "CloudParams | None" = None,
                                       samples_per_pixel: int = 64):
        """EXTENSIONS over several GPUs: the strips `rank` owns of a spectral (kind="spectral") or cloud (kind="cloud")
        frame, every pixel stored at its final position in each framebuffer of `dst_ptrs`.  Returns the stats dict."""
        cam = camera.to_c()
        stats = A.ft_render_stats()
        arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
        st = C.c_void_p(stream) if stream else None
        if kind == "spectral":
            A.check(self.lib.ft_render_spectral_gather_device(self.handle, C.byref(cam), int(size.X), int(size.Y), C.c_float(epsilon),
                                                              C.c_float(length), int(n_wavelengths), int(max_bounces), C.c_uint64(seed),
                                                              int(strip_width), int(rank), int(world), arr, len(dst_ptrs), st,
                                                              C.byref(stats)))
        elif kind == "cloud":
            cp = (cloud or CloudParams()).to_c()
            A.check(self.lib.ft_render_cloud_gather_device(self.handle, C.byref(cam), int(size.X), int(size.Y), C.c_float(epsilon),
                                                           C.c_float(length), C.byref(cp), int(samples_per_pixel), C.c_uint64(seed),
                                                           int(strip_width), int(rank), int(world), arr, len(dst_ptrs), st,
                                                           C.byref(stats)))
        else:
            raise ValueError("kind must be 'spectral' or 'cloud'")
        return stats.as_dict()

    def trace_rays(self, rays: np.ndarray, flags: int = 0):
        """Batched SdfScene.trace: rays float32 [n,8] = origin, direction, length, epsilon."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        rgb = np.empty((n, 3), np.float32)
        leaf = np.empty(n, np.int32)
        steps = np.empty(n, np.int32)
        shadow = np.empty(n, np.uint32)
        stats = A.ft_render_stats()
        A.check(self.lib.ft_trace_rays(self.handle, n, _ptr(rays), int(flags), _ptr(rgb), _ptr(leaf),
                                       _ptr(steps), _ptr(shadow), C.byref(stats)))
        return dict(rgb=rgb, leaf=leaf, steps=steps, shadow=shadow), stats.as_dict()

    def distance(self, points: np.ndarray, flags: int = 0) -> np.ndarray:
        points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        out = np.empty(points.shape[0], np.float32)
        A.check(self.lib.ft_distance(self.handle, points.shape[0], _ptr(points), int(flags), _ptr(out)))
        return out

    def __call__(self, ray: Ray):
        """the reference's `Ray -> FColor`"""
        res, _ = self.trace_rays(ray.as_row()[None, :])
        return tuple(float(x) for x in res["rgb"][0])


class CloudParams:
    """EXTENSION: parameters of the cloud medium plus the two tables of the reference's dead-code lattice
    noise (Noise.fs): `gradients` is Noise.gradientTable regenerated exactly as Noise.fs:31-33,72-76 draws it
    (System.Random(666): 256 value-table draws first, then z and theta per entry); `perm` plays the role of
    Noise.permutationTable (Noise.fs:7-23) -- the reference's literal is not copied here: the default is a Fisher-Yates
    shuffle by the same generator, and an F# host can pass the reference's own table."""

    def __init__(self, density=6.0, frequency=1.5, step=0.05, fade=0.25, bias=0.45, gain=0.9, perm=None, gradients=None):
        self.density, self.frequency, self.step, self.fade, self.bias, self.gain = (float(density), float(frequency),
                                                                                   float(step), float(fade), float(bias),
                                                                                   float(gain))
        if perm is None or gradients is None:
            dperm, dgrad = CloudParams.default_tables()
        self.perm = np.ascontiguousarray(dperm if perm is None else perm, dtype=np.int32)
        self.gradients = np.ascontiguousarray(dgrad if gradients is None else gradients, dtype=np.float64).reshape(256, 3)
        assert self.perm.shape == (256,)

    _tables = None

    @staticmethod
    def default_tables():
        if CloudParams._tables is None:
            import math
            from .dotnet_random import DotNetRandom
            rng = DotNetRandom(666)  # Noise.fs:31
            for _ in range(256):  # Noise.fs:33 valueTable draws come first
                rng.NextDouble()
            grad = np.empty((256, 3), np.float64)
            for i in range(256):  # Noise.fs:72-76
                z = rng.NextDouble() * 2.0 - 1.0
                r = math.sqrt(1.0 - z * z)
                theta = 2.0 * math.pi * rng.NextDouble()
                grad[i] = (r * math.cos(theta), r * math.sin(theta), z)
            perm = np.arange(256, dtype=np.int32)
            for i in range(255, 0, -1):  # Fisher-Yates with the same generator
                j = int(rng.NextDouble() * (i + 1))
                perm[i], perm[j] = perm[j], perm[i]
            CloudParams._tables = (perm, grad)
        return CloudParams._tables

    def to_c(self) -> "A.ft_cloud_params":
        cp = A.ft_cloud_params()
        cp.density, cp.frequency, cp.step, cp.fade, cp.bias, cp.gain = (self.density, self.frequency, self.step,
                                                                        self.fade, self.bias, self.gain)
        cp.perm = self.perm.ctypes.data_as(C.POINTER(C.c_int32))
        cp.gradients = self.gradients.ctypes.data_as(C.POINTER(C.c_double))
        return cp


class Image:
    """Image.fs:25-35"""

    @staticmethod
    def render(epsilon: float, length: float, imageSize: ImageSize, camera: Camera, trace) -> np.ndarray:
        if not isinstance(trace, SceneTracer):
            raise TypeError("Image.render needs the SceneTracer returned by SdfScene.trace; arbitrary "
                            "Ray -> FColor closures cannot run on the device (FT_ERR_UNSUPPORTED_NODE)")
        out, _, _ = trace.render(camera, imageSize, epsilon, length)
        return out

    @staticmethod
    def toBitmap(gamma: float, seed: int, image: np.ndarray, device: int = 0) -> np.ndarray:
        """Image.toColors gamma rng image |> Image.toBitmap (Image.fs:37-72) on the GPU: uint8 [Y, X, 3], top-down
        B,G,R scanlines in the reference's pixel order.  `seed` keys the deterministic Philox dither that replaces
        the reference's shared (racy) System.Random."""
        image = np.ascontiguousarray(image, dtype=np.float32)
        X, Y = image.shape[0], image.shape[1]
        out = np.zeros((Y, X, 3), np.uint8)
        A.check(A.load_library().ft_to_bitmap(int(device), _ptr(image), X, Y, C.c_float(gamma), C.c_uint64(seed),
                                              _ptr(out), 3 * X))
        return out

    @staticmethod
    def saveBitmap(path: str, bitmap: np.ndarray):
        """Image.saveBitmap path (Image.fs:87-90): 24-bit BMP from the scanlines of Image.toBitmap"""
        bitmap = np.ascontiguousarray(bitmap, dtype=np.uint8)
        Y, X = bitmap.shape[0], bitmap.shape[1]
        A.check(A.load_library().ft_save_bitmap(str(path).encode(), _ptr(bitmap), X, Y, 3 * X))
