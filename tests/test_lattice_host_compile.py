"""CPU tier, lattice geometry (configs C3 / C4): engine sources vs oracle.
Cell / lattice indices and crossing tokens bit-exact, one generation identical."""
import numpy as np
import pytest

import cases
from openmc_b200 import lib
from oracle.binding import Oracle


@pytest.mark.parametrize("name", ["assembly", "c5g7", "pwr_core"])
def test_lattice_geometry_bit_exact(emu_lib, name):
    model, box, of = cases.build_lattice(name)
    o = Oracle(model)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 32768)
    xyz, uvw = cases.geometry_probe_points(box, 30000)
    for a, b in zip(eng.find_cell(xyz, uvw), o.find_cell(xyz, uvw)):
        assert np.array_equal(a, b)
    dg, sg, lg, tg = eng.distance_to_boundary(xyz, uvw)
    do, so, lo_, to = o.distance_to_boundary(xyz, uvw)
    assert np.array_equal(sg, so) and np.array_equal(lg, lo_) and np.array_equal(tg, to)
    assert np.array_equal(dg, do)
    assert int(model["n_coord_levels"]) == (2 if name == "assembly" else 3)
    eng.finalize()


@pytest.mark.parametrize("threshold", ["0", "100000"])
@pytest.mark.parametrize("name", ["assembly", "c5g7", "pwr_core"])
def test_lattice_generation_matches_oracle(emu_lib, name, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model, box, of = cases.build_lattice(name)
    o = Oracle(model)
    n = 1200
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 2048)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0 and out_o.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    assert out.n_crossings == out_o.n_crossings and out.n_collisions == out_o.n_collisions
    for f in ("r", "u", "E", "time", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-13)
    m = int(model["tally_result_ofs"][1])
    v = np.zeros(m); eng.tally_results(0, v)
    np.testing.assert_allclose(v, o.tally_results(0, m), rtol=1e-11, atol=1e-300)
    eng.finalize()


def test_many_nuclide_lookup_bit_exact(emu_lib):
    """config C5 fuel (40 nuclides here): the lockstep lookup kernels (batches of 4 nuclides)
    give the grid indices and the sums of the one-at-a-time walk, bit for bit"""
    model, box, of = cases.build_lattice("pwr_core")
    n_nuc = np.diff(np.asarray(model["mat_nuc_ofs"]))
    mat = int(np.argmax(n_nuc))
    assert n_nuc[mat] >= 16
    o = Oracle(model)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 1024)
    E = cases.xs_probe_energies(model, None, 4000)
    mats = np.full(E.size, mat, np.int32)
    xs_e, ig_e = eng.calculate_xs(E, mats, int(n_nuc[mat]))
    xs_o, ig_o = o.calculate_xs(E, mats, int(n_nuc[mat]))
    assert np.array_equal(ig_e, ig_o) and np.array_equal(xs_e, xs_o)
    eng.finalize()


@pytest.mark.parametrize("env", [
    {"OMCGPU_SORT_COLL": "1"},                             # collision list in (material, energy) order too
    {"OMCGPU_SORT_XS": "0"},                               # fused, index-ordered sweeps with the lockstep lookup
    {"OMCGPU_SORT_XS": "0", "OMCGPU_SWEEP": "1"},          # merged sweep, fused lookup
    {"OMCGPU_ORDERED": "0"},                               # lists appended with atomics
    {"OMCGPU_LOW_COUNT": "100000"},                        # lockstep lookup for every material
    {"OMCGPU_BATCH_XS": "0", "OMCGPU_SORT_XS": "0"},       # one nuclide at a time
])
def test_many_nuclide_scheduling_variants_bank_the_same_sites(emu_lib, env, monkeypatch):
    """config C5 (40-nuclide fuel here): every optional path of the many-nuclide machinery
    leaves the histories' arithmetic alone -- identical events and fission bank"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "64")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    model, box, of = cases.build_lattice("pwr_core")
    o = Oracle(model)
    n = 800
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 1024)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    for f in ("r", "u", "E", "wgt", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    eng.finalize()


@pytest.mark.parametrize("env", [{}, {"OMCGPU_XS_PREFIX": "0"}, {"OMCGPU_SORT_XS": "0"}])
def test_nuclide_sampling_from_block_prefix_sums(emu_lib, env, monkeypatch):
    """a 100-nuclide fuel: the macro lookup leaves the running sum of rho sigma_t after every 32
    nuclides in the particle's prefix row, and the collision's sample_nuclide walk
    (src/physics.cpp:494-518) starts at the block the cutoff falls into -- the nuclide it finds,
    hence every later event and the fission bank, are those of the walk from nuclide 0 (oracle)"""
    from openmc_b200 import examples, nuclide
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "64")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    m, libx, box, of = examples.pwr_core_ce(n_grid=300, n_fuel_nuclides=100, tallies=True)
    model = nuclide.build_ce_model(m, libx)
    o = Oracle(model)
    n = 600
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 1024)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    for f in ("r", "u", "E", "wgt", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    eng.finalize()


@pytest.mark.parametrize("sweep", ["1", "2"])
@pytest.mark.parametrize("threshold", ["0", "100000"])
@pytest.mark.parametrize("name", ["hex_bundle_y", "hex_bundle_x3d"])
def test_hexagonal_lattices(emu_lib, name, threshold, sweep, monkeypatch):
    """HexLattice (src/lattice.cpp:458-1009: get_indices, get_local_position, distance with the
    neighbour-tile frames, are_valid_indices, the skewed square storage and its distribcell
    offsets), both orientations, 2-D and 3-D: one generation of the engine sources equals the
    oracle -- which equals the reference's own binary bit for bit on these two models
    (tests/test_reference_end_to_end.py)"""
    import reference_runs as rr
    from openmc_b200 import lib
    from oracle.binding import Oracle
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)
    m, libx, model, box, of, ce = rr.build(name)
    o = Oracle(model)
    n = 1500
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0 and out_o.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    assert out.n_crossings == out_o.n_crossings
    for f in ("r", "u", "E", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-13)
    for t in range(int(model["n_tallies"])):
        k = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(k); eng.tally_results(t, v)
        vo = o.tally_results(t, k)
        assert vo.sum() > 0
        np.testing.assert_allclose(v, vo, rtol=1e-11, atol=1e-300)
    eng.finalize()
