#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ (run in the container that has
/root/reference; the fixtures are committed because the reference does not
travel to the GPU box).

Sources:
* ref_helpers.npz -- outputs of the reference's OWN translation units
  (src/random_lcg.cpp, src/math_functions.cpp, src/random_dist.cpp,
  include/openmc/search.h) compiled verbatim into oracle/_ref by oracle/Makefile.
* kat.json -- the literal known answers of the reference's unit tests
  (tests/cpp_unit_tests/test_math.cpp:74-96, 204-277) and the regression goldens
  (tests/regression_tests/mg_*/results_true.dat).
* ce_lookup.npz / geometry.npz -- oracle outputs on fixed inputs for the CE
  lookup (grid indices, cross sections) and the C2 pincell geometry, so that the
  GPU tier has a fixture that does not depend on the oracle being rebuilt.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import binding  # noqa: E402
import cases  # noqa: E402


def helpers():
    ref = binding.Reference().lib
    g = np.random.default_rng(20261017)
    n = 256
    seeds = g.integers(0, 2**63, n, dtype=np.uint64)
    skip = np.concatenate([g.integers(0, 2**62, n - 8, dtype=np.uint64),
                           np.array([0, 1, 2, 152917, 2**63, 2**64 - 1, 7, 10**12], np.uint64)])
    fut = np.array([ref.ref_future_seed(int(a), int(b)) for a, b in zip(skip, seeds)], np.uint64)
    draws = np.zeros((n, 6))
    for i in range(n):
        s = C.c_uint64(int(seeds[i]))
        for j in range(6):
            draws[i, j] = ref.ref_prn(C.byref(s))
    init = np.array([ref.ref_init_seed(int(i), int(o)) for i in (1, 17, 10**9) for o in (0, 1, 2, 3)],
                    np.uint64)
    # Legendre, rotate_angle, spectra
    leg_c = g.normal(size=(64, 8)); leg_n = g.integers(0, 8, 64); leg_x = g.uniform(-1, 1, 64)
    leg = np.array([ref.ref_evaluate_legendre(int(leg_n[i]), leg_c[i].ctypes.data_as(C.POINTER(C.c_double)),
                                              C.c_double(leg_x[i])) for i in range(64)])
    u = g.normal(size=(64, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    u[:4] = [0.0, 0.0, 1.0]
    mu = g.uniform(-1, 1, 64)
    rot = np.zeros((64, 3)); rot_seed = np.zeros(64, np.uint64)
    for i in range(64):
        v = u[i].copy(); s = C.c_uint64(int(seeds[i]))
        ref.ref_rotate_angle(v.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(mu[i]), None, C.byref(s))
        rot[i] = v; rot_seed[i] = s.value
    watt = np.zeros(64); maxw = np.zeros(64)
    for i in range(64):
        s = C.c_uint64(int(seeds[i]))
        watt[i] = ref.ref_watt_spectrum(C.c_double(0.988e6), C.c_double(2.249e-6), C.byref(s))
        maxw[i] = ref.ref_maxwell_spectrum(C.c_double(1.3e6), C.byref(s))
    grid = np.array([1.0, 2.0, 2.0, 3.0, 5.0, 8.0])
    vals = np.array([1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0, 7.9, 8.0])
    lbi = np.array([ref.ref_lower_bound_index(grid.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(6),
                                              C.c_double(v)) for v in vals], np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_helpers.npz"), seeds=seeds, skip=skip, future_seed=fut,
                        prn=draws, init_seed=init, leg_c=leg_c, leg_n=leg_n, leg_x=leg_x, leg=leg,
                        rot_u=u, rot_mu=mu, rot=rot, rot_seed=rot_seed, watt=watt, maxwell=maxw,
                        lbi_grid=grid, lbi_vals=vals, lbi=lbi)


def kat():
    d = {
        "source": "tests/cpp_unit_tests/test_math.cpp:74-96,204-277; tests/regression_tests/mg_*/results_true.dat",
        "evaluate_legendre": {"coeffs": [1.0] * 11, "x": [-1.0, -0.5, 0.0, 0.5, 1.0],
                              "ref": [5.5, -0.45597649, -1.35351562, -2.7730999, 60.5]},
        "rotate_angle_no_phi": {"uvw0": [1.0, 0.0, 0.0], "mu": 0.9, "seed": 1,
                                "ref": [0.9, -0.422746750548505, 0.10623175090659095]},
        "maxwell_spectrum": {"T": 0.5, "seed": 1, "ref": 0.27767406743161277},
        "watt_spectrum": {"a": 0.5, "b": 0.75, "seed": 1, "ref": 0.30957476387766697},
        "regression": {k: {"k_combined": v[0], "k_std": v[1], "tallies": v[2]} for k, v in cases.GOLDEN.items()},
    }
    json.dump(d, open(os.path.join(HERE, "kat.json"), "w"), indent=1)


def oracle_vectors():
    """The flattened model itself is part of the fixture (ce_model.npz): the
    synthetic library is generated with numpy transcendental functions whose last
    bit may differ between machines, and the fixture must not depend on that."""
    model, box, of, libx = cases.build_ce(n_grid=1000)
    cases.save_model_npz(os.path.join(HERE, "ce_model.npz"), model)
    model = cases.load_model_npz(os.path.join(HERE, "ce_model.npz"))
    o = binding.Oracle(model)
    E = cases.xs_probe_energies(model, libx, 3000)
    out = {"E": E}
    for mat in range(int(model["n_materials"])):
        xs, ig = o.calculate_xs(E, np.full(E.size, mat, np.int32), 4)
        out[f"xs{mat}"] = xs; out[f"ig{mat}"] = ig
    np.savez_compressed(os.path.join(HERE, "ce_lookup.npz"), **out)
    g = np.random.default_rng(4)
    n = 20000
    box = np.asarray(box, float)
    lo, hi = box[:3] - 0.05 * (box[3:] - box[:3]), box[3:] + 0.05 * (box[3:] - box[:3])
    xyz = lo + g.uniform(size=(n, 3)) * (hi - lo)
    uvw = g.normal(size=(n, 3)); uvw /= np.linalg.norm(uvw, axis=1)[:, None]
    cell, inst, mat, nco = o.find_cell(xyz, uvw)
    d, s, lv, lt = o.distance_to_boundary(xyz, uvw)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), xyz=xyz, uvw=uvw, cell=cell, instance=inst,
                        material=mat, n_coord=nco, distance=d, surface=s, level=lv, lattice_trans=lt)


if __name__ == "__main__":
    helpers()
    kat()
    oracle_vectors()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith((".npz", ".json"))))
