// ft_math.h -- float32 vector arithmetic in the exact operation order of System.Numerics, as the
// reference's F# code uses it ("Arithmetic conventions").  Shared by the host
// scene compiler and the device kernels.  One IEEE rounding per operation: the translation units that
// include this for the bit-exact path are compiled with -fmad=false (device) / -ffp-contract=off (host).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FT_HD __host__ __device__ __forceinline__
#else
#define FT_HD inline
#endif

namespace ftm {

struct V3 {
    float x, y, z;
};

FT_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
FT_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
FT_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
FT_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
FT_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
FT_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
FT_HD V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
FT_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }

// Vector3.Dot: x64 dpps reduction (x*x' + y*y') + z*z'
FT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
FT_HD float length2(V3 a) { return dot(a, a); }
FT_HD float length(V3 a) { return sqrtf(dot(a, a)); }
FT_HD float distance(V3 a, V3 b) { return length(a - b); }
FT_HD float distance2(V3 a, V3 b) { return length2(a - b); }
FT_HD V3 normalize(V3 a) { return a / length(a); }
FT_HD V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
FT_HD V3 vmin(V3 a, V3 b) { return V3{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
FT_HD V3 vmax(V3 a, V3 b) { return V3{a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
// Vector3.Lerp (.NET 5+): a*(1-t) + b*t
FT_HD V3 lerp(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
// Math.fs:69  Vector3.inverseLength v = v / v.LengthSquared()
FT_HD V3 inverse_length(V3 a) { return a / length2(a); }

// MathF.Floor |> int on x64 (cvttss2si): out-of-range and NaN give INT_MIN (Math.fs:57)
FT_HD int floor_to_int(float x) {
    float f = floorf(x);
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
FT_HD int ceiling_to_int(float x) {
    float f = ceilf(x);
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
// MathI.clamp lo hi x = Math.Max(lo, Math.Min(hi, x)) (Math.fs:23)
FT_HD int clamp_int(int lo, int hi, int x) {
    int m = hi < x ? hi : x;
    return lo > m ? lo : m;
}

}  // namespace ftm
