"""A SECOND, independent restatement of the reference's whole ray-test -- SdfScene.trace with everything under it -- in
plain Python over numpy float32 scalars, written from the F# source (paths relative to /root/reference/src/FrayTracer) and
sharing no code with oracle/ft_oracle.c or the CUDA path.  It covers what BASELINE config 1 needs (Program.fs): torus and
sphere primitives, union (with buildSpatialLookup), intersect, subtract, SdfObject.union's material arg-min, directional and
point lights.  Slow by design; tests compare small frames bit for bit (tests/test_independent_restatement.py).

Arithmetic: every operation on np.float32 scalars rounds once to binary32, in source order; Vector3.Dot is
(x*x' + y*y') + z*z'; Normalize and `/ float32` divide per component (.NET 6).  Array.sortInPlaceBy is restated as a stable
sort: tests only use scenes whose LowerBounds have no ties inside a cell (asserted)."""
import math

import numpy as np

f32 = np.float32
INF = f32(np.inf)


def v(x, y, z):
    return (f32(x), f32(y), f32(z))


def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def mul(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
def scale(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def divs(a, s): return (a[0] / s, a[1] / s, a[2] / s)
def dot(a, b): return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
def length(a): return np.sqrt(dot(a, a))
def distance(a, b): return length(sub(a, b))
def normalize(a): return divs(a, length(a))
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def net_max(a, b): return a if a > b else b  # MathF.Max for non-NaN, non-signed-zero-sensitive use
def net_min(a, b): return a if a < b else b


# ---- SdfForm.Primitive (SdfForm.fs:125-135, 181-203): returns (distance function, boundary (centre, radius))
def sphere(center, radius):
    c, r = v(*center), f32(radius)
    return (lambda p: distance(c, p) - r), (c, r)


def torus(center, normal, major, minor):
    c, n = v(*center), normalize(v(*normal))  # :183-186
    major, minor = f32(major), f32(minor)
    plane_d = -dot(c, n)  # :189

    def dist(p):
        dp = dot(p, n) + plane_d  # :190
        dc = distance(c, sub(p, scale(n, dp)))  # :191  position - (distanceToPlane * data.Normal)
        q = dc - major  # :192
        return np.sqrt(dp * dp + q * q) - minor  # :194 Vector2.Length

    return dist, (c, major + minor)  # :199-202


# ---- SdfBoundary.buildSpatialLookup (SdfBoundary.fs:225-282)
class Lookup:
    def __init__(self, bounds):
        n = len(bounds)
        mins = [sub(c, (r, r, r)) for c, r in bounds]  # AABB.getMin / getMax
        maxs = [add(c, (r, r, r)) for c, r in bounds]
        amin, amax = mins[0], maxs[0]
        for m in mins[1:]:
            amin = tuple(net_min(a, b) for a, b in zip(amin, m))
        for m in maxs[1:]:
            amax = tuple(net_max(a, b) for a, b in zip(amax, m))
        total = bounds[0][1]
        for _, r in bounds[1:]:
            total = total + r  # Seq.reduce (+)
        count_size = f32(1.5) * (total / f32(n))  # :233
        size = sub(amax, amin)
        cnt = max(1, int(math.ceil(float(size[0] / count_size))))  # :237-239 (aabbSize.X for all three axes)
        self.n = cnt
        self.amin = amin
        self.cell = divs(size, f32(cnt))  # :240 (Vector3 / Vector3 with equal components)
        self.inv = (f32(1) / self.cell[0], f32(1) / self.cell[1], f32(1) / self.cell[2])  # :241
        self.bounds = bounds
        self.cells = {}

    def cell_of(self, p):  # :276-282
        c = mul(sub(p, self.amin), self.inv)
        return tuple(min(self.n - 1, max(0, int(math.floor(float(x))))) for x in c)

    def get(self, p):
        key = self.cell_of(p)
        if key not in self.cells:  # built lazily: the same values Array3D.Parallel.init computes eagerly
            half = scale(self.cell, f32(0.5))
            center = add(add(self.amin, half), mul(self.cell, v(*key)))  # :246
            upper = min(distance(c, center) + r for c, r in self.bounds) + length(half)  # :248-253
            items = [(distance(c, center) - r, i) for i, (c, r) in enumerate(self.bounds)]
            items = [it for it in items if it[0] < upper]  # :258
            items.sort(key=lambda t: t[0])  # :267-268 (stable; no ties allowed, see module docstring)
            lbs = [float(t[0]) for t in items]
            assert len(set(lbs)) == len(lbs), "tie between LowerBounds: the reference's unstable sort decides, not this restatement"
            self.cells[key] = (center, items)
        return self.cells[key]


def union_form(prims):  # SdfForm.fs:14-40
    dists = [d for d, _ in prims]
    look = Lookup([b for _, b in prims])

    def dist(p):
        center, items = look.get(p)
        dtc = distance(p, center)  # :25
        m = dists[items[0][1]](p)  # :26
        for lb, i in items[1:]:
            c, r = look.bounds[i]
            if m > lb - dtc and m > distance(c, p) - r:  # :29-31
                m = net_min(dists[i](p), m)  # :33  MathF.min min
        return m

    return dist, look


def union_material_argmin(prims, look, p):  # SdfObject.fs:27-46 (the second lookup is identical by construction)
    dists = [d for d, _ in prims]
    center, items = look.get(p)
    best = items[0][1]
    m = dists[best](p)
    dtc = distance(p, center)
    for lb, i in items:  # starts again at item 0
        c, r = look.bounds[i]
        if m > lb - dtc and m > distance(c, p) - r:
            d = dists[i](p)
            if d < m:
                m, best = d, i
    return best


def intersect_form(forms):  # SdfForm.fs:51-67, forms = [(dist, boundary)]
    def dist(p):
        m = forms[0][0](p)
        for d, (c, r) in forms[1:]:
            if m < distance(c, p) + r:  # getMaxDistance
                m = net_max(d(p), m)
        return m
    return dist


def subtract_form(a, b):  # SdfForm.fs:42-49
    return lambda p: net_max(-(b(p)), a(p))


# ---- SdfForm.tryTrace / normalFromRay, SdfObject.tryTrace, SdfLight, SdfScene.trace
def try_trace(dist, o, d, length_, eps):  # SdfForm.fs:93-104 -> (hit?, origin, length, steps)
    steps = 0
    while True:
        if length_ <= 0:
            return False, o, length_, steps
        dd = dist(o)
        steps += 1
        if dd < eps:
            return True, o, length_, steps
        o = add(o, scale(d, dd))  # Ray.move
        length_ = length_ - dd


def normal_from_ray(dist, o, d, eps):  # SdfForm.fs:106-115
    p = add(o, scale(d, -eps))
    h = eps * f32(0.125)
    n = (dist((p[0] + h, p[1], p[2])), dist((p[0], p[1] + h, p[2])), dist((p[0], p[1], p[2] + h)))
    d0 = dist(p)
    return normalize(sub(n, (d0, d0, d0)))


class Scene:
    """Program.fs:67-83: subtract(intersect(union tori, [sphere]), sphere); `lights` = [("dir", v, rgb) | ("point", v, rgb)]"""

    def __init__(self, tori, colors, clip_sphere, cut_sphere, lights, background):
        self.prims = [torus(*t) for t in tori]
        self.colors = [v(*c) for c in colors]
        u, self.look = union_form(self.prims)
        self.dist = u
        if clip_sphere is not None:
            self.dist = intersect_form([(self.dist, None), sphere(*clip_sphere)])
        if cut_sphere is not None:
            self.dist = subtract_form(self.dist, sphere(*cut_sphere)[0])
        self.lights = [(k, normalize(scale(v(*x), f32(-1))) if k == "dir" else v(*x), v(*rgb)) for k, x, rgb in lights]  # SdfLight.fs:7
        self.bg = v(*background)

    def object_try_trace(self, o, d, length_, eps, shade):  # SdfObject.fs:66-78
        hit, ho, hl, steps = try_trace(self.dist, o, d, length_, eps)
        if not hit:
            return None
        if not shade:
            return True
        n = normal_from_ray(self.dist, ho, d, eps)
        leaf = union_material_argmin(self.prims, self.look, ho)  # Material.Color at the UN-moved hit origin
        back = add(ho, scale(d, -eps))  # Ray.move -epsilon
        return back, n, leaf, steps

    def trace(self, o, d, length_, eps):  # SdfScene.fs:7-28 -> (rgb, leaf, primary steps)
        r = self.object_try_trace(o, d, length_, eps, True)
        if r is None:
            return self.bg, -1
        pos, n, leaf, _ = r
        lc = self.bg
        for kind, x, rgb in self.lights:
            ldir = x if kind == "dir" else normalize(sub(x, pos))
            cos = dot(n, ldir)
            if cos > 0:
                if kind == "dir":  # SdfLight.fs:10-20
                    blocked = self.object_try_trace(pos, x, f32(1000), eps, False)
                    intensity = rgb
                else:  # SdfLight.fs:26-41
                    diff = sub(x, pos)
                    d2 = dot(diff, diff)
                    blocked = self.object_try_trace(pos, divs(diff, d2), np.sqrt(d2), eps, False)
                    intensity = divs(rgb, d2)
                if blocked is None:
                    lc = add(lc, scale(intensity, cos))
        pi_inv = f32(1) / f32(math.pi)  # Math.fs:30
        return mul(self.colors[leaf], scale(lc, pi_inv)), leaf


def render(scene, cam_pos, forward, right_scaled, up_scaled, X, Y, eps, length_, pixels=None):
    """Image.render (Image.fs:17-35, Camera.fs:44-54) for the given pixels (default: all) -> {(x, y): (rgb, leaf)}"""
    m = f32(max(X, Y))
    out = {}
    for x, y in (pixels if pixels is not None else [(x, y) for x in range(X) for y in range(Y)]):
        u, w = f32(x) / m, f32(y) / m
        d = normalize(add(add(forward, scale(right_scaled, u - f32(0.5))), scale(up_scaled, w - f32(0.5))))
        out[(x, y)] = scene.trace(cam_pos, d, f32(length_), f32(eps))
    return out


def look_at(position, target, up, fov):  # Camera.fs:10-14, 33-42
    nps = f32(math.sin(float(f32(fov) * f32(0.5))))  # sin of degrees-as-radians, evaluated in double, rounded once
    pos = v(*position)
    forward = normalize(sub(v(*target), pos))
    right = normalize(cross(v(*up), forward))
    return pos, forward, scale(right, nps), scale(cross(forward, right), nps)
