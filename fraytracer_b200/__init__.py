"""fraytracer_b200 -- B200-native (sm_100a) per-ray hot path of JanWosnitza/FrayTracer.

The product is `csrc/libfraytracer_b200.so` (CUDA kernels + scene compiler behind the C ABI of
include/fraytracer_b200.h).  This package is the thin host-side mirror of the reference's F#
API over that ABI.  There is no CPU fallback.
"""
from ._abi import (FT_FLAG_COUNTERS, FT_FLAG_DEFAULT, FT_FLAG_FAST_MATH, FrayTracerError, load_library)
from . import distributed  # noqa: F401
from .api import (Camera, CloudParams, FColor, Image, ImageSize, Lens, Ray, SceneDescription, SceneTracer, SdfForm,
                  SdfLight, SdfMaterial, SdfObject, SdfScene)

__all__ = [
    "Camera", "CloudParams", "FColor", "Image", "ImageSize", "Lens", "Ray", "SceneDescription", "SceneTracer", "SdfForm",
    "SdfLight", "SdfMaterial", "SdfObject", "SdfScene", "FrayTracerError", "load_library",
    "FT_FLAG_DEFAULT", "FT_FLAG_COUNTERS", "FT_FLAG_FAST_MATH",
]
