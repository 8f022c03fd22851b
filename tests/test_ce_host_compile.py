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


def check_collide_against_reference_vectors(make_engine):
    """the DEVICE function collision_ce (csrc/device.h) through the C ABI (omcgpu_collide) against
    tests/golden/refcoll_ce.npz -- what the reference's own collision() (src/physics.cpp compiled
    unmodified, tests/golden/make_reference_collisions.py) left of the same neutrons: energy,
    direction, weight, k-absorption score, seed and banked fission sites bit for bit, analog
    absorption and survival biasing"""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    model = cases.load_model_npz(os.path.join(golden, "ce_model.npz"))
    v = np.load(os.path.join(golden, "refcoll_ce.npz"))
    mat, E, u, seeds = (np.ascontiguousarray(v[k]) for k in ("material", "E_in", "u_in", "seed_in"))
    n = E.size
    for tag, survival in (("a", 0), ("s", 1)):
        model.struct.survival_biasing = survival
        eng = make_engine()
        eng.init(model, 0, max(n, 1024))
        r = eng.collide(mat, E, u, seeds, float(v["k_eff"]), float(v[f"{tag}_wgt0"]))
        eng.finalize()
        got = dict(E=r["E"], u=r["u"], wgt=r["wgt"], k_abs=r["k_abs"], seed=r["seed"],
                   site_ofs=np.concatenate([[0], np.cumsum(r["n_sites"])]).astype(np.int64),
                   site_E=r["sites"]["E"], site_u=r["sites"]["u"],
                   site_group=r["sites"]["delayed_group"], site_time=r["sites"]["time"])
        # the sorted bank lists the sites of neutron k (parent_id k + 1) in progeny order
        assert np.array_equal(r["sites"]["parent_id"], np.repeat(np.arange(1, n + 1), r["n_sites"]))
        for name in got:
            assert np.array_equal(got[name], v[f"{tag}_{name}"]), (tag, name)


def test_device_collision_equals_reference_vectors(emu_lib):
    check_collide_against_reference_vectors(lambda: lib.Engine(emu_lib))
