"""The restated helpers against the reference's OWN translation units
(random_lcg.cpp, math_functions.cpp, random_dist.cpp, search.h, position.h)
compiled verbatim into oracle/_ref by oracle/Makefile."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding
from openmc_b200 import rng

pytestmark = pytest.mark.skipif(
    not (os.path.isdir("/root/reference/src") or os.path.exists(binding.REF_LIB)),
    reason="reference tree / oracle/_ref not available")


@pytest.fixture(scope="module")
def ref():
    return binding.Reference().lib


@pytest.fixture(scope="module")
def orc():
    binding.build()
    L = C.CDLL(binding.LIB)
    L.orc_evaluate_legendre.restype = C.c_double
    L.orc_maxwell_spectrum.restype = C.c_double
    L.orc_watt_spectrum.restype = C.c_double
    L.orc_lower_bound_index.restype = C.c_int64
    return L


def test_prn_and_future_seed(ref, orc):
    g = np.random.default_rng(1)
    seeds = g.integers(0, 2**63, 200, dtype=np.uint64)
    ns = np.concatenate([g.integers(0, 2**62, 190, dtype=np.uint64),
                         np.array([0, 1, 2, 152917, 2**63, 2**64 - 1, 7, 8, 9, 10], np.uint64)])
    out = np.zeros(200, np.uint64)
    orc.orc_future_seed(ns.ctypes.data_as(C.POINTER(C.c_uint64)),
                        seeds.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(200),
                        out.ctypes.data_as(C.POINTER(C.c_uint64)))
    for i in range(200):
        want = ref.ref_future_seed(int(ns[i]), int(seeds[i]))
        assert int(out[i]) == want
        assert rng.future_seed(int(ns[i]), int(seeds[i])) == want
    assert np.array_equal(rng.future_seed_v(ns, seeds), out)
    draws = np.zeros((200, 5))
    orc.orc_prn(seeds.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(200), C.c_int32(5),
                draws.ctypes.data_as(C.POINTER(C.c_double)))
    for i in range(200):
        s = C.c_uint64(int(seeds[i]))
        sp = int(seeds[i])
        for j in range(5):
            want = ref.ref_prn(C.byref(s))
            assert draws[i, j] == want
            xi, sp = rng.prn(sp)
            assert xi == want and sp == s.value
    xi, s2 = rng.prn_v(seeds)
    assert np.array_equal(xi, draws[:, 0])
    # init_seed with the default master seed / stride
    for ident in (1, 17, 10**9):
        assert rng.init_seed(ident, 1) == ref.ref_init_seed(ident, 1)


def test_math_helpers(ref, orc):
    g = np.random.default_rng(2)
    for _ in range(200):
        n = int(g.integers(0, 8))
        data = g.normal(size=n + 1)
        x = float(g.uniform(-1, 1))
        p = data.ctypes.data_as(C.POINTER(C.c_double))
        assert orc.orc_evaluate_legendre(n, p, C.c_double(x)) == \
            ref.ref_evaluate_legendre(n, p, C.c_double(x))
    for _ in range(200):
        u = g.normal(size=3); u /= np.linalg.norm(u)
        if g.uniform() < 0.1:
            u = np.array([0.0, 0.0, 1.0])
        mu = float(g.uniform(-1, 1))
        seed = int(g.integers(0, 2**63))
        a = u.copy(); b = u.copy()
        sa = C.c_uint64(seed); sb = C.c_uint64(seed)
        orc.orc_rotate_angle(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(mu), None, C.byref(sa))
        ref.ref_rotate_angle(b.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(mu), None, C.byref(sb))
        assert np.array_equal(a, b) and sa.value == sb.value
    for _ in range(100):
        seed = int(g.integers(0, 2**63))
        sa = C.c_uint64(seed); sb = C.c_uint64(seed)
        assert orc.orc_watt_spectrum(C.c_double(0.988e6), C.c_double(2.249e-6), C.byref(sa)) == \
            ref.ref_watt_spectrum(C.c_double(0.988e6), C.c_double(2.249e-6), C.byref(sb))
        assert orc.orc_maxwell_spectrum(C.c_double(1.3e6), C.byref(sa)) == \
            ref.ref_maxwell_spectrum(C.c_double(1.3e6), C.byref(sb))
        assert sa.value == sb.value


def test_lower_bound_index(ref, orc):
    # include/openmc/search.h:14-20: k-1 when value equals grid point k>0, 0 at the first point
    a = np.array([1.0, 2.0, 2.0, 3.0, 5.0, 8.0])
    p = a.ctypes.data_as(C.POINTER(C.c_double))
    for v in (1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0, 7.9, 8.0):
        assert orc.orc_lower_bound_index(p, C.c_int64(a.size), C.c_double(v)) == \
            ref.ref_lower_bound_index(p, C.c_int64(a.size), C.c_double(v))
