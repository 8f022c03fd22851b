"""Committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py
from the reference's own translation units and unit-test known answers) against
the oracle, the host-compiled engine sources and -- on the GPU tier -- the CUDA
library.  Nothing here reads /root/reference."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import cases
from openmc_b200 import lib, rng
from oracle import binding
from oracle.binding import Oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def helpers():
    return np.load(os.path.join(G, "ref_helpers.npz"))


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(G, "kat.json")))


@pytest.fixture(scope="module")
def orc():
    binding.build()
    L = C.CDLL(binding.LIB)
    for f in ("orc_evaluate_legendre", "orc_maxwell_spectrum", "orc_watt_spectrum"):
        getattr(L, f).restype = C.c_double
    L.orc_lower_bound_index.restype = C.c_int64
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- oracle / host logic vs the reference's outputs (CPU tier) -----------------
def test_oracle_rng_vs_reference_vectors(helpers):
    o = Oracle(cases.build("mg_legendre")[0])
    assert np.array_equal(o.future_seed(helpers["skip"], helpers["seeds"]), helpers["future_seed"])
    assert np.array_equal(o.prn(helpers["seeds"], 6), helpers["prn"])          # bit-exact doubles
    assert np.array_equal(rng.future_seed_v(helpers["skip"], helpers["seeds"]), helpers["future_seed"])
    k = 0
    for ident in (1, 17, 10**9):
        for off in (0, 1, 2, 3):
            assert rng.init_seed(ident, off) == int(helpers["init_seed"][k])
            k += 1


def test_oracle_math_vs_reference_vectors(helpers, orc):
    for i in range(64):
        c = np.ascontiguousarray(helpers["leg_c"][i])
        assert orc.orc_evaluate_legendre(int(helpers["leg_n"][i]), _pd(c),
                                         C.c_double(helpers["leg_x"][i])) == helpers["leg"][i]
        u = helpers["rot_u"][i].copy()
        s = C.c_uint64(int(helpers["seeds"][i]))
        orc.orc_rotate_angle(_pd(u), C.c_double(helpers["rot_mu"][i]), None, C.byref(s))
        assert np.array_equal(u, helpers["rot"][i]) and s.value == int(helpers["rot_seed"][i])
        s = C.c_uint64(int(helpers["seeds"][i]))
        assert orc.orc_watt_spectrum(C.c_double(0.988e6), C.c_double(2.249e-6), C.byref(s)) == helpers["watt"][i]
        assert orc.orc_maxwell_spectrum(C.c_double(1.3e6), C.byref(s)) == helpers["maxwell"][i]
    grid = np.ascontiguousarray(helpers["lbi_grid"])
    for v, want in zip(helpers["lbi_vals"], helpers["lbi"]):
        assert orc.orc_lower_bound_index(_pd(grid), C.c_int64(grid.size), C.c_double(v)) == want


def test_reference_unit_test_known_answers(kat, orc):
    """tests/cpp_unit_tests/test_math.cpp:74-96, 204-277"""
    k = kat["evaluate_legendre"]
    c = np.array(k["coeffs"])
    got = [orc.orc_evaluate_legendre(len(c) - 1, _pd(c), C.c_double(x)) for x in k["x"]]
    np.testing.assert_allclose(got, k["ref"], rtol=1e-7)
    k = kat["rotate_angle_no_phi"]
    u = np.array(k["uvw0"]); s = C.c_uint64(k["seed"])
    orc.orc_rotate_angle(_pd(u), C.c_double(k["mu"]), None, C.byref(s))
    np.testing.assert_allclose(u, k["ref"], rtol=1e-12)
    s = C.c_uint64(kat["maxwell_spectrum"]["seed"])
    assert orc.orc_maxwell_spectrum(C.c_double(kat["maxwell_spectrum"]["T"]), C.byref(s)) == \
        kat["maxwell_spectrum"]["ref"]                                            # REQUIRE(==) in the reference
    s = C.c_uint64(kat["watt_spectrum"]["seed"])
    assert orc.orc_watt_spectrum(C.c_double(kat["watt_spectrum"]["a"]), C.c_double(kat["watt_spectrum"]["b"]),
                                 C.byref(s)) == kat["watt_spectrum"]["ref"]


def _check_lookup(eng):
    model = cases.load_model_npz(os.path.join(G, "ce_model.npz"))
    v = np.load(os.path.join(G, "ce_lookup.npz"))
    eng.init(model, 0, 1024)
    for mat in range(int(model["n_materials"])):
        xs, ig = eng.calculate_xs(v["E"], np.full(v["E"].size, mat, np.int32), 4)
        assert np.array_equal(ig, v[f"ig{mat}"])                                  # grid indices: bit-exact
        np.testing.assert_allclose(xs, v[f"xs{mat}"], rtol=1e-12, atol=0)         # north_star tolerance
    eng.finalize()


def _check_geometry(eng):
    model = cases.load_model_npz(os.path.join(G, "ce_model.npz"))
    v = np.load(os.path.join(G, "geometry.npz"))
    eng.init(model, 0, 32768)
    cell, inst, mat, nco = eng.find_cell(v["xyz"], v["uvw"])
    assert np.array_equal(cell, v["cell"]) and np.array_equal(inst, v["instance"])
    assert np.array_equal(mat, v["material"]) and np.array_equal(nco, v["n_coord"])
    d, s, lv, lt = eng.distance_to_boundary(v["xyz"], v["uvw"])
    assert np.array_equal(s, v["surface"]) and np.array_equal(lv, v["level"])
    assert np.array_equal(lt, v["lattice_trans"])
    np.testing.assert_allclose(d, v["distance"], rtol=1e-12)
    eng.finalize()


def test_engine_sources_vs_fixtures(emu_lib):
    _check_lookup(lib.Engine(emu_lib))
    _check_geometry(lib.Engine(emu_lib))


# ---- CUDA library vs the fixtures (GPU tier) -----------------------------------
@pytest.mark.gpu
def test_gpu_rng_vs_reference_vectors(gpu_engine, helpers):
    assert np.array_equal(gpu_engine.future_seed(helpers["skip"], helpers["seeds"]), helpers["future_seed"])
    assert np.array_equal(gpu_engine.prn(helpers["seeds"], 6), helpers["prn"])


@pytest.mark.gpu
def test_gpu_lookup_and_geometry_vs_fixtures(gpu_engine):
    _check_lookup(gpu_engine)
    _check_geometry(gpu_engine)
