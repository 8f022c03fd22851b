"""CPU tier, continuous energy: engine sources (host-compiled for tests) vs oracle."""
import numpy as np
import pytest

import cases
from openmc_b200 import lib
from oracle.binding import Oracle


def test_macro_xs_and_grid_index(emu_lib):
    model, box, of, libx = cases.build_ce()
    o = Oracle(model)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 1024)
    E = cases.xs_probe_energies(model, libx, 20000)
    for mat in range(int(model["n_materials"])):
        mats = np.full(E.size, mat, np.int32)
        xs_e, ig_e = eng.calculate_xs(E, mats, 4)
        xs_o, ig_o = o.calculate_xs(E, mats, 4)
        assert np.array_equal(ig_e, ig_o)          # energy-grid indices: bit-exact
        assert np.array_equal(xs_e, xs_o)          # same libm on the host: identical
    eng.finalize()


@pytest.mark.parametrize("threshold", ["0", "100000"])
def test_generation_matches_oracle(emu_lib, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model, box, of, _ = cases.build_ce()
    o = Oracle(model)
    n = 1500
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 2048)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0 and out_o.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    for f in ("r", "u", "E", "time", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-13)
    for t in range(2):
        m = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(m); eng.tally_results(t, v)
        np.testing.assert_allclose(v, o.tally_results(t, m), rtol=1e-11, atol=1e-300)
    eng.finalize()


def test_eigenvalue_matches_oracle(emu_lib):
    model, box, of, _ = cases.build_ce()
    o = Oracle(model)
    src = o.sample_initial_source(800, box, of)
    r = o.run_eigenvalue(src, 6, 2)
    sim = lib.Simulation(lib.Engine(emu_lib), model, 800, 6, 2, src)
    k = sim.run()
    assert abs(k[0] - r["k_combined"][0]) < 1e-12 and abs(k[1] - r["k_combined"][1]) < 1e-12
    np.testing.assert_allclose(np.array(sim.k_generation), r["k_generation"], rtol=1e-13)
    sim.simulation_finalize()


def test_device_collision_equals_reference_vectors(emu_lib):
    """the DEVICE function collision_ce (csrc/device.h, host-compiled; the emu-only doorway
    emu_collide_ce runs it once per neutron) against tests/golden/refcoll_ce.npz -- what the
    reference's own collision() (src/physics.cpp compiled unmodified, make_reference_collisions.py)
    left of the same neutrons: energy, direction, weight, k-absorption score, seed and banked
    fission sites bit for bit, analog absorption and survival biasing"""
    import ctypes as C
    import os
    import ref_geometry
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    model = cases.load_model_npz(os.path.join(golden, "ce_model.npz"))
    v = np.load(os.path.join(golden, "refcoll_ce.npz"))
    mat, E, u, seeds = (np.ascontiguousarray(v[k]) for k in ("material", "E_in", "u_in", "seed_in"))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = E.size
    for tag, survival in (("a", 0), ("s", 1)):
        model.struct.survival_biasing = survival
        eng = lib.Engine(emu_lib)
        eng.init(model, 0, 1024)
        b, cap = ref_geometry._collision_buffers(n)
        eng.lib.emu_collide_ce.restype = C.c_int64
        tot = eng.lib.emu_collide_ce(C.c_int64(n), p(mat), p(E), p(u), p(seeds), C.c_double(float(v["k_eff"])),
                                     C.c_double(float(v[f"{tag}_wgt0"])), p(b["E"]), p(b["u"]), p(b["wgt"]),
                                     p(b["k_abs"]), p(b["seed"]), p(b["site_ofs"]), p(b["site_E"]), p(b["site_u"]),
                                     p(b["site_group"]), p(b["site_time"]), C.c_int64(cap))
        eng.finalize()
        assert tot >= 0, tot
        b = ref_geometry._trim_sites(b, tot)
        for name in ref_geometry.COLLISION_FIELDS:
            assert np.array_equal(b[name], v[f"{tag}_{name}"]), (tag, name)
