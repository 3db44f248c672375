"""System.Random(seed) of .NET (seeded constructor: Knuth's subtractive generator) and the
reference's extension methods on it (src/FrayTracer/Random.fs:9-40).

Needed only to regenerate the reference's synthetic scenes (Program.fs:14 uses
System.Random(19)); nothing on the device path draws random numbers.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
_MBIG = 2147483647
_MSEED = 161803398


def _wrap32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


class DotNetRandom:
    def __init__(self, seed: int):
        sub = _MBIG if seed == -(1 << 31) else abs(seed)
        sa = [0] * 56
        mj = _MSEED - sub
        sa[55] = mj
        mk = 1
        ii = 0
        for _ in range(1, 55):
            ii += 21
            if ii >= 55:
                ii -= 55
            sa[ii] = mk
            mk = mj - mk
            if mk < 0:
                mk += _MBIG
            mj = sa[ii]
        for _ in range(4):
            for i in range(1, 56):
                n = i + 30
                if n >= 55:
                    n -= 55
                sa[i] = _wrap32(sa[i] - sa[1 + n])
                if sa[i] < 0:
                    sa[i] += _MBIG
        self._sa = sa
        self._inext = 0
        self._inextp = 21

    def Next(self) -> int:
        a, b = self._inext + 1, self._inextp + 1
        if a >= 56:
            a = 1
        if b >= 56:
            b = 1
        v = _wrap32(self._sa[a] - self._sa[b])
        if v == _MBIG:
            v -= 1
        if v < 0:
            v += _MBIG
        self._sa[a] = v
        self._inext, self._inextp = a, b
        return v

    def NextDouble(self) -> float:
        return self.Next() * (1.0 / _MBIG)

    # ---- Random.fs extension members (float32 arithmetic, one rounding per operation)
    def range_01(self) -> np.float32:  # Random.fs:9
        return f32(self.NextDouble())

    def range(self, lo, hi) -> np.float32:  # Random.fs:11
        lo, hi = f32(lo), f32(hi)
        return f32(lo + f32(self.range_01() * f32(hi - lo)))

    @staticmethod
    def _len2(v) -> np.float32:  # Vector3.LengthSquared = Dot(v,v) = (x*x + y*y) + z*z
        return f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))

    def pointInBall(self, radius) -> np.ndarray:  # Random.fs:27-32
        radius = f32(radius)
        while True:
            v = np.array([self.range(-1, 1), self.range(-1, 1), self.range(-1, 1)], dtype=f32)
            if self._len2(v) <= f32(1.0):
                return (v * radius).astype(f32)

    def pointOnSphere(self, radius) -> np.ndarray:  # Random.fs:34-40
        radius = f32(radius)
        while True:
            v = np.array([self.range(-1, 1), self.range(-1, 1), self.range(-1, 1)], dtype=f32)
            l2 = self._len2(v)
            if f32(0.01) <= l2 <= f32(1.0):
                return ((v / f32(np.sqrt(l2))).astype(f32) * radius).astype(f32)
