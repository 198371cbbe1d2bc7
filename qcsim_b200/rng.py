"""std::mt19937_64 + the draw `1. - uniform_real_distribution<double>(0,1)(rng)` the reference
uses for every measurement (QubitRegister.h:171,210,267,621,707, members :723-724), restated so
that a register seeded with the same 64-bit value produces the same draws as libstdc++.
"""
from __future__ import annotations

import math
import time

_M64 = (1 << 64) - 1


class Mt19937_64:
    NN, MM = 312, 156
    MATRIX_A = 0xB5026F5AA96619E9
    UM, LM = 0xFFFFFFFF80000000, 0x7FFFFFFF

    def __init__(self, seed: int = 5489):
        self.mt = [0] * self.NN
        self.mti = self.NN
        self.seed(seed)

    def seed(self, seed: int) -> None:
        mt = self.mt
        mt[0] = seed & _M64
        for i in range(1, self.NN):
            mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & _M64
        self.mti = self.NN

    def _refill(self) -> None:
        mt, NN, MM = self.mt, self.NN, self.MM
        for i in range(NN):
            x = (mt[i] & self.UM) | (mt[(i + 1) % NN] & self.LM)
            mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (self.MATRIX_A if x & 1 else 0)
        self.mti = 0

    def __call__(self) -> int:
        if self.mti >= self.NN:
            self._refill()
        x = self.mt[self.mti]
        self.mti += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & _M64

    def uniform01(self) -> float:
        """std::generate_canonical<double, 53> for a 64-bit engine (one engine call)."""
        u = float(self()) / 18446744073709551616.0
        if u >= 1.0:
            u = math.nextafter(1.0, 0.0)
        return u

    def draw(self) -> float:
        """`1. - uniformZeroOne(rng)`: in (0, 1], excludes 0 as a probability."""
        return 1.0 - self.uniform01()


def time_seed(addseed: int = 0) -> int:
    """Clock-based seed in the spirit of QubitRegister.h:26-34 (not reproducible by design)."""
    if addseed == 0:
        import secrets

        addseed = secrets.randbits(32)
    return (time.time_ns() + addseed) & _M64
