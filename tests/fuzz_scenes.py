"""random scene generator for the fuzz tests: random composite trees over random primitives"""
import numpy as np

from fraytracer_b200 import SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene

P = SdfForm.Primitive


def random_primitive(rng, spread=2.5):
    c = rng.normal(size=3) * spread / 2
    kind = rng.integers(0, 5)
    if kind == 0:
        return P.sphere(Center=c, Radius=rng.uniform(0.2, 0.9))
    if kind == 1:
        d = rng.normal(size=3)
        return P.capsule(From=c, To=c + d / np.linalg.norm(d) * rng.uniform(0.3, 1.5), Radius=rng.uniform(0.1, 0.4))
    if kind == 2:
        return P.torus(Center=c, Normal=rng.normal(size=3), MajorRadius=rng.uniform(0.2, 0.7), MinorRadius=rng.uniform(0.08, 0.3))
    if kind == 4:  # EXTENSION primitive (no reference): capped cone
        d = rng.normal(size=3)
        return P.cone(From=c, To=c + d / np.linalg.norm(d) * rng.uniform(0.3, 1.5), FromRadius=rng.uniform(0.0, 0.6),
                      ToRadius=rng.uniform(0.0, 0.6))
    a, b = rng.normal(size=3), rng.normal(size=3)
    return P.triangle(V1=c, V2=c + a / np.linalg.norm(a) * rng.uniform(0.3, 0.9),
                      V3=c + b / np.linalg.norm(b) * rng.uniform(0.3, 0.9), Radius=rng.uniform(0.08, 0.3))


def random_form(rng, depth):
    if depth == 0 or rng.random() < 0.35:
        return random_primitive(rng)
    op = rng.integers(0, 4)
    if op == 0:
        return SdfForm.union([random_form(rng, depth - 1) for _ in range(rng.integers(2, 7))])
    if op == 1:
        return SdfForm.subtract(random_form(rng, depth - 1), random_form(rng, depth - 1))
    if op == 2:
        return SdfForm.intersect([random_form(rng, depth - 1) for _ in range(rng.integers(2, 4))])
    return SdfForm.unionSmooth(float(rng.uniform(0.05, 0.4)), [random_form(rng, depth - 1) for _ in range(rng.integers(2, 4))])


def has_smooth(form):
    from fraytracer_b200 import _abi as A

    return form.kind == A.FT_FORM_UNION_SMOOTH or any(has_smooth(c) for c in form.children)


def random_object(rng, depth, allow_smooth=False):
    def leaf():
        while True:
            f = random_form(rng, min(depth, 2)) if rng.random() < 0.3 else random_primitive(rng)
            if allow_smooth or not has_smooth(f):
                return SdfObject.create(SdfMaterial.createSolid(rng.uniform(0.1, 1.0, size=3)), f)
    if depth == 0:
        return leaf()
    op = rng.integers(0, 4)
    if op <= 1:
        return SdfObject.union([random_object(rng, depth - 1, allow_smooth) if rng.random() < 0.3 else leaf()
                                for _ in range(rng.integers(2, 12))])
    if op == 2:
        return SdfObject.subtract(random_object(rng, depth - 1, allow_smooth), random_primitive(rng))
    return SdfObject.intersect(random_object(rng, depth - 1, allow_smooth),
                               [P.sphere(Center=rng.normal(size=3) * 0.3, Radius=rng.uniform(1.5, 3.5))])


def random_scene(seed, allow_smooth=False):
    rng = np.random.default_rng(seed)
    obj = random_object(rng, depth=int(rng.integers(1, 4)), allow_smooth=allow_smooth)
    lights = [SdfLight.directional(rng.normal(size=3), rng.uniform(0.2, 1.0, size=3))]
    if rng.random() < 0.7:
        lights.append(SdfLight.point(rng.normal(size=3) * 3, rng.uniform(1, 10, size=3)))
    return SdfScene(Object=obj, BackgroundColor=tuple(rng.uniform(0, 0.3, size=3)), Lights=lights)
