"""Continuous-energy lookup against the REFERENCE ITSELF: Nuclide::init_grid and
Nuclide::calculate_xs of src/nuclide.cpp, compiled unmodified into
oracle/_ref/libref_geom.so, on the synthetic pointwise tables.  The hash table, every grid
index (incl. energies ON grid points, one ulp either side, and on hash-bin edges) and the
density-weighted sums must be identical in the oracle; the CUDA path is held to the
committed reference vectors (indices bit-exact, cross sections <= 1e-12: north_star)."""
import os

import numpy as np
import pytest

import cases
import ref_geometry
from oracle.binding import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not os.path.exists(ref_geometry.LIB),
                               reason="oracle/_ref/libref_geom.so not built (no /root/reference here)")


@needs_ref
def test_hash_table_and_lookup_equal_reference():
    model, box, of, libx = cases.build_ce(n_grid=4000)
    ref = ref_geometry.ReferenceXS(model)
    gi = np.asarray(model["nuc_grid_index"]).reshape(ref.nn, ref.n_log_bins + 1)
    for k in range(ref.nn):
        assert np.array_equal(ref.grid_index(k), gi[k])        # Nuclide::init_grid
    E = ref.in_range(cases.xs_probe_energies(model, libx, 30000))
    o = Oracle(model)
    for mat in range(int(model["n_materials"])):
        n_nuc = int(model["mat_nuc_ofs"][mat + 1] - model["mat_nuc_ofs"][mat])
        if n_nuc == 0:
            continue
        xs_r, ig_r = ref.macro(mat, E)
        xs_o, ig_o = o.calculate_xs(E, np.full(E.size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_o[:, :n_nuc], ig_r)
        assert np.array_equal(xs_o, xs_r)


def _fixture():
    model = cases.load_model_npz(os.path.join(GOLDEN, "ce_model.npz"))
    return model, np.load(os.path.join(GOLDEN, "refxs_ce.npz"))


def test_committed_reference_vectors_are_reproduced_by_the_oracle():
    model, v = _fixture()
    o = Oracle(model)
    for mat in range(int(model["n_materials"])):
        if f"xs{mat}" not in v.files:
            continue
        n_nuc = v[f"ig{mat}"].shape[1]
        xs_o, ig_o = o.calculate_xs(v["E"], np.full(v["E"].size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_o[:, :n_nuc], v[f"ig{mat}"]) and np.array_equal(xs_o, v[f"xs{mat}"])


@pytest.mark.gpu
def test_gpu_lookup_equals_reference_vectors(gpu_engine):
    model, v = _fixture()
    gpu_engine.init(model, 0, 1024)
    for mat in range(int(model["n_materials"])):
        if f"xs{mat}" not in v.files:
            continue
        n_nuc = v[f"ig{mat}"].shape[1]
        xs_g, ig_g = gpu_engine.calculate_xs(v["E"], np.full(v["E"].size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_g[:, :n_nuc], v[f"ig{mat}"])                  # bit-exact
        np.testing.assert_allclose(xs_g, v[f"xs{mat}"], rtol=1e-12, atol=0)     # north_star tolerance
    gpu_engine.finalize()
