"""Synthetic scenes of BASELINE.json's configs, generated exactly like the reference's
console program does (src/FrayTracer.Console/Program.fs:14-93): same System.Random stream,
same draw order, same float32 arithmetic.

config 1  default_scene()           Program.fs verbatim: 1000 tori, intersect sphere 3.5, subtract sphere 2.5
config 2  spheres_capsules_scene()  512 randomSphere + 512 randomCapsule in one union (capsule stands in
                                    for the reference's non-existent cone), 1920x1080
config 3  spectral_glass_scene()    EXTENSION (no reference): config-2 generator, every 3rd object dispersive glass
config 4  cloud_scene()             EXTENSION (no reference): solids + large spheres of cloud medium, 64 spp
config 5  mixed_scene()             256 each torus/sphere/capsule/triangle in one union, 7680x4320
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .api import (Camera, FColor, ImageSize, Lens, SdfForm, SdfLight, SdfMaterial, SdfObject, SdfScene)
from .dotnet_random import DotNetRandom

f32 = np.float32


class SceneBuilder:
    """the `random*` helpers of Program.fs:28-65 over one System.Random stream"""

    def __init__(self, seed: int = 19):
        self.rng = DotNetRandom(seed)  # Program.fs:14

    def randomMaterial(self) -> SdfMaterial:  # Program.fs:28-30
        r = self.rng.range_01()
        g = self.rng.range_01()
        b = self.rng.range_01()
        return SdfMaterial.createSolid(FColor.ofRGB(r, g, b))

    def randomSphere(self) -> SdfObject:  # Program.fs:32-37
        center = self.rng.pointInBall(4.0)
        radius = self.rng.range(0.3, 1.0)
        form = SdfForm.Primitive.sphere(Center=center, Radius=radius)
        return SdfObject.create(self.randomMaterial(), form)

    def randomCapsule(self) -> SdfObject:  # Program.fs:39-46
        center = self.rng.pointInBall(4.0)
        length = self.rng.range(0.5, 2.0)
        to = (center + self.rng.pointOnSphere(length)).astype(f32)
        radius = self.rng.range(0.1, 0.3)
        form = SdfForm.Primitive.capsule(From=center, To=to, Radius=radius)
        return SdfObject.create(self.randomMaterial(), form)

    def randomCone(self) -> SdfObject:
        """EXTENSION (the reference has no cone): randomCapsule's draws (Program.fs:39-46) with the tip radius 0"""
        center = self.rng.pointInBall(4.0)
        length = self.rng.range(0.5, 2.0)
        to = (center + self.rng.pointOnSphere(length)).astype(f32)
        radius = self.rng.range(0.1, 0.3)
        form = SdfForm.Primitive.cone(From=center, To=to, FromRadius=float(radius) * 2.0, ToRadius=0.0)
        return SdfObject.create(self.randomMaterial(), form)

    def randomTorus(self) -> SdfObject:  # Program.fs:48-55
        center = self.rng.pointInBall(4.0)
        normal = self.rng.pointOnSphere(1.0)
        major = self.rng.range(0.1, 0.4)
        minor = self.rng.range(0.1, 0.3)
        form = SdfForm.Primitive.torus(Center=center, Normal=normal, MajorRadius=major, MinorRadius=minor)
        return SdfObject.create(self.randomMaterial(), form)

    def randomTriangle(self) -> SdfObject:  # Program.fs:57-65
        v1 = self.rng.pointInBall(4.0)
        l2 = self.rng.range(0.2, 0.6)
        v2 = (v1 + self.rng.pointOnSphere(l2)).astype(f32)
        l3 = self.rng.range(0.2, 0.6)
        v3 = (v1 + self.rng.pointOnSphere(l3)).astype(f32)
        radius = self.rng.range(0.1, 0.3)
        form = SdfForm.Primitive.triangle(V1=v1, V2=v2, V3=v3, Radius=radius)
        return SdfObject.create(self.randomMaterial(), form)


def default_lights():  # Program.fs:79-82
    return [
        SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(0.5, 0.5, 0.5)),
        SdfLight.point((-0.5, 0.0, -2.0), FColor.ofRGB(10.0, 0.0, 0.0)),
    ]


# Program.fs:16-22 -- every workload uses the reference's camera; kept as plain numbers so that the oracle can build the same
# camera through its own Lens.create / Camera.lookAt
CAMERA_SPEC = dict(Position=(0.0, 0.0, -10.0), LookAt=(0.0, 0.0, 0.0), Up=(0.0, 1.0, 0.0), FieldOfView=60.0)


def default_camera() -> Camera:  # Program.fs:16-22
    return Camera.lookAt(Position=CAMERA_SPEC["Position"], LookAt=CAMERA_SPEC["LookAt"], Up=CAMERA_SPEC["Up"],
                         Lens=Lens.create(CAMERA_SPEC["FieldOfView"]))


@dataclass
class Workload:
    name: str
    scene: SdfScene
    camera: Camera
    size: ImageSize
    epsilon: float
    length: float
    leaves: list  # leaf SdfObjects in generation order


def default_scene(seed: int = 19, n_tori: int = 1000, size: int = 1000) -> Workload:
    """config 1 -- Program.fs:67-93"""
    b = SceneBuilder(seed)
    tori = [b.randomTorus() for _ in range(n_tori)]
    obj = SdfObject.subtract(
        SdfObject.intersect(
            SdfObject.union(tori),
            [SdfForm.Primitive.sphere(Center=(0.0, 0.0, 0.0), Radius=3.5)]),
        SdfForm.Primitive.sphere(Center=(-0.5, 1.0, -2.0), Radius=2.5))
    scene = SdfScene(Object=obj, BackgroundColor=FColor.ofRGB(0.1, 0.1, 0.1), Lights=default_lights())
    return Workload("config1_default_scene", scene, default_camera(), ImageSize(size, size), 0.01, 30.0, tori)


def spheres_capsules_scene(seed: int = 19, n_spheres: int = 512, n_capsules: int = 512,
                           width: int = 1920, height: int = 1080) -> Workload:
    """config 2 -- SURVEY.md section 8(d)"""
    b = SceneBuilder(seed)
    objs = [b.randomSphere() for _ in range(n_spheres)] + [b.randomCapsule() for _ in range(n_capsules)]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.1, 0.1, 0.1),
                     Lights=default_lights())
    return Workload("config2_spheres_capsules_1080p", scene, default_camera(), ImageSize(width, height),
                    0.01, 30.0, objs)


def spheres_cones_scene(seed: int = 19, n_spheres: int = 512, n_cones: int = 512, width: int = 1920,
                        height: int = 1080) -> Workload:
    """config 2 as BASELINE.json words it ("spheres + cones") -- EXTENSION: the cone primitive has no reference
    implementation, so this scene is a parity case against the repository's own CPU restatement;
    the bench workload stays `spheres_capsules_scene`, whose primitives the reference defines."""
    b = SceneBuilder(seed)
    objs = [b.randomSphere() for _ in range(n_spheres)] + [b.randomCone() for _ in range(n_cones)]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.1, 0.1, 0.1),
                     Lights=default_lights())
    return Workload("config2b_spheres_cones_1080p", scene, default_camera(), ImageSize(width, height), 0.01, 30.0, objs)


def mixed_scene(seed: int = 19, n_each: int = 256, width: int = 7680, height: int = 4320) -> Workload:
    """config 5 -- SURVEY.md section 8(d)"""
    b = SceneBuilder(seed)
    objs = ([b.randomTorus() for _ in range(n_each)] + [b.randomSphere() for _ in range(n_each)]
            + [b.randomCapsule() for _ in range(n_each)] + [b.randomTriangle() for _ in range(n_each)])
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.1, 0.1, 0.1),
                     Lights=default_lights())
    return Workload("config5_mixed_1024_8k", scene, default_camera(), ImageSize(width, height), 0.01, 30.0, objs)


# BK7-like Cauchy coefficients with the dispersion term exaggerated 10x so that the colour split is visible at
# image scale (n = 1.5046 + 0.042 / lambda^2, lambda in micrometres)
GLASS_CAUCHY = (1.5046, 0.042)


def spectral_glass_scene(seed: int = 19, n_spheres: int = 192, n_capsules: int = 192, glass_every: int = 3,
                         width: int = 3840, height: int = 2160) -> Workload:
    """config 3 -- EXTENSION, no reference implementation: the config-2 generator
    (Program.fs:32-46 helpers, same System.Random stream) where every `glass_every`-th object is dispersive glass,
    lit by a white directional and a white point light."""
    b = SceneBuilder(seed)
    objs = [b.randomSphere() for _ in range(n_spheres)] + [b.randomCapsule() for _ in range(n_capsules)]
    glass = SdfMaterial.createGlass(*GLASS_CAUCHY)
    objs = [SdfObject.create(glass, o.form) if i % glass_every == 0 else o for i, o in enumerate(objs)]
    lights = [SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0)),
              SdfLight.point((-0.5, 0.0, -2.0), FColor.ofRGB(10.0, 10.0, 10.0))]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.3, 0.4, 0.6), Lights=lights)
    return Workload("config3_spectral_glass_4k", scene, default_camera(), ImageSize(width, height), 0.01, 30.0, objs)


def _bound(form: SdfForm):
    """bounding sphere of a sphere / capsule form (SdfForm.fs:131-134, 166-169), float64 is fine for placement"""
    p = form.params
    if form.kind == 1:  # FT_FORM_SPHERE
        return np.array(p[0:3]), p[3]
    a, b = np.array(p[0:3]), np.array(p[3:6])
    return (a + b) * 0.5, p[6] + float(np.linalg.norm(b - a)) * 0.5


def cloud_scene(seed: int = 19, n_spheres: int = 64, n_capsules: int = 64, n_clouds: int = 24,
                width: int = 1920, height: int = 1080) -> Workload:
    """config 4 -- EXTENSION, no reference implementation: solid spheres and capsules
    from the config-2 generator (Program.fs:32-46 helpers) plus `n_clouds` large spheres of cloud medium drawn from the
    same System.Random stream (rejected while they touch a solid: what overlaps a cloud is hidden inside it), lit by a
    white directional and a white point light."""
    b = SceneBuilder(seed)
    objs = [b.randomSphere() for _ in range(n_spheres)] + [b.randomCapsule() for _ in range(n_capsules)]
    bounds = [_bound(o.form) for o in objs]
    cloud = SdfMaterial.createCloud(FColor.ofRGB(0.95, 0.95, 0.95))
    placed = 0
    for _ in range(100000):
        if placed == n_clouds:
            break
        center = b.rng.pointInBall(5.0)
        radius = float(b.rng.range(0.6, 1.6))
        if any(np.linalg.norm(np.asarray(center, np.float64) - c) < radius + r + 0.05 for c, r in bounds):
            continue
        objs.append(SdfObject.create(cloud, SdfForm.Primitive.sphere(Center=center, Radius=radius)))
        placed += 1
    lights = [SdfLight.directional((-0.5, -1.0, 1.0), FColor.ofRGB(2.0, 2.0, 2.0)),
              SdfLight.point((-0.5, 0.0, -2.0), FColor.ofRGB(10.0, 10.0, 10.0))]
    scene = SdfScene(Object=SdfObject.union(objs), BackgroundColor=FColor.ofRGB(0.3, 0.4, 0.6), Lights=lights)
    return Workload("config4_cloud_1080p_64spp", scene, default_camera(), ImageSize(width, height), 0.01, 30.0, objs)
