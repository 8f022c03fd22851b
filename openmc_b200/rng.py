"""Host-side random numbers: src/random_lcg.cpp (prn :32-44, future_seed :91-120,
init_seed :60-64), scalar (python ints) and vectorised (numpy uint64)."""
from __future__ import annotations

import numpy as np

PRN_MULT = 6364136223846793005
PRN_ADD = 1442695040888963407
M64 = (1 << 64) - 1
STREAM_TRACKING, STREAM_SOURCE = 0, 1


def prn(seed: int):
    """returns (xi, new_seed)"""
    seed = (PRN_MULT * seed + PRN_ADD) & M64
    word = (((seed >> ((seed >> 59) + 5)) ^ seed) * 12605985483714917081) & M64
    result = (word >> 43) ^ word
    return float(np.ldexp(np.float64(result), -64)), seed


def future_seed(n: int, seed: int) -> int:
    g, c, g_new, c_new = PRN_MULT, PRN_ADD, 1, 0
    n &= M64
    while n > 0:
        if n & 1:
            g_new = (g_new * g) & M64
            c_new = (c_new * g + c) & M64
        c = (c * (g + 1)) & M64
        g = (g * g) & M64
        n >>= 1
    return (g_new * seed + c_new) & M64


def init_seed(id_: int, offset: int, master_seed: int = 1, stride: int = 152917) -> int:
    return future_seed((id_ * stride) & M64, (master_seed + offset) & M64)


# ---- vectorised --------------------------------------------------------------
def future_seed_v(n: np.ndarray, seed: np.ndarray) -> np.ndarray:
    n = n.astype(np.uint64).copy()
    seed = seed.astype(np.uint64)
    with np.errstate(over="ignore"):
        g = np.full(n.shape, PRN_MULT, np.uint64)
        c = np.full(n.shape, PRN_ADD, np.uint64)
        g_new = np.ones(n.shape, np.uint64)
        c_new = np.zeros(n.shape, np.uint64)
        one = np.uint64(1)
        while np.any(n > 0):
            odd = (n & one).astype(bool)
            g_new = np.where(odd, g_new * g, g_new)
            c_new = np.where(odd, c_new * g + c, c_new)
            c = c * (g + one)
            g = g * g
            n >>= one
        return g_new * seed + c_new


def prn_v(seed: np.ndarray):
    """vectorised prn: returns (xi, new_seed) for uint64 arrays"""
    with np.errstate(over="ignore"):
        seed = seed * np.uint64(PRN_MULT) + np.uint64(PRN_ADD)
        shift = (seed >> np.uint64(59)) + np.uint64(5)
        word = ((seed >> shift) ^ seed) * np.uint64(12605985483714917081)
        result = (word >> np.uint64(43)) ^ word
    return np.ldexp(result.astype(np.float64), -64), seed
