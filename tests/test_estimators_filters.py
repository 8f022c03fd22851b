"""Collision-estimator tallies (score_collision_tally, src/tallies/tally_scoring.cpp:2543-2620) and the
DistribcellFilter (src/tallies/filter_distribcell.cpp:40-63): engine sources (host-compiled, CPU
tier; CUDA in the GPU tier) against the oracle, and the flattener's distribcell offset tables
against the reference's own prepare_distribcell (src/geometry_aux.cpp:327-446)."""
import os

import numpy as np
import pytest

from openmc_b200 import examples, lib, mgxs, nuclide
from openmc_b200.model import Filter, Mesh, Tally
from oracle.binding import Oracle


def assembly_with_estimators(n_grid=1500):
    """config C3 geometry with: pin mesh flux/fission by collision estimator, a spectrum by
    collision estimator, distribcell (every fuel pin) flux by tracklength and by collision"""
    m, libx, box, of = examples.assembly_ce(n_grid=n_grid)
    pitch = 21.42
    fuel_cell = m.cells[0].id          # first cell of the fuel-pin universe
    m.meshes = [Mesh(1, [17, 17], [-pitch / 2, -pitch / 2], [pitch / 2, pitch / 2])]
    m.filters = [Filter(1, "mesh", 1), Filter(2, "energy", list(np.logspace(-5, 7.3, 13))),
                 Filter(3, "distribcell", [fuel_cell])]
    m.tallies = [Tally(1, [1], ["flux", "fission"], "collision"),
                 Tally(2, [2], ["flux", "total", "absorption"], "collision"),
                 Tally(3, [3], ["flux", "nu-fission"], "tracklength"),
                 Tally(4, [3, 2], ["flux", "scatter"], "collision"),
                 Tally(5, [1], ["flux"], "tracklength")]
    return nuclide.build_ce_model(m, libx), box, of


def c5g7_with_estimators():
    m, libx, box, of = examples.c5g7_core(mesh_dim=17)
    pin_fuel = m.cells[0].id
    edges = sorted(libx.group_edges)
    m.filters = list(m.filters) + [Filter(50, "distribcell", [pin_fuel]), Filter(51, "energy", list(edges))]
    m.tallies = list(m.tallies) + [Tally(50, [50], ["flux", "fission"], "collision"),
                                   Tally(51, [51], ["flux", "scatter", "nu-fission"], "collision"),
                                   Tally(52, [m.filters[0].id], ["flux", "total"], "collision")]
    return mgxs.build_mg_model(m, libx), box, of


def _tallies_of(model, get):
    out = []
    for t in range(int(model["n_tallies"])):
        n = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        out.append(get(t, n))
    return out


def _check_generation(eng, build, n, rtol, exact_bank=True):
    model, box, of = build()
    o = Oracle(model)
    src = o.sample_initial_source(n, box, of)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    if exact_bank:
        assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
        assert np.array_equal(fb["parent_id"], fb_o["parent_id"])

    def dev(t, k):
        v = np.zeros(k)
        eng.tally_results(t, v)
        return v
    got = _tallies_of(model, dev)
    want = _tallies_of(model, o.tally_results)
    for t, (a, b) in enumerate(zip(got, want)):
        assert b.sum() > 0, t                       # every tally of the model is exercised
        if exact_bank:
            np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-300, err_msg=f"tally {t}")
        else:
            assert abs(a.sum() - b.sum()) <= 2e-4 * b.sum(), t
    eng.finalize()
    return model


@pytest.mark.parametrize("threshold", ["0", "100000"])
def test_collision_estimator_and_distribcell_ce(emu_lib, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model = _check_generation(lib.Engine(emu_lib), assembly_with_estimators, 1500, 1e-11)
    # the distribcell filter has one bin per fuel pin of the 17 x 17 lattice
    assert int(model["filter_n_bins"][2]) == 17 * 17 - 25
    assert int(model["n_collision_tallies"]) == 0      # derived by omcgpu_init, not by the flattener


@pytest.mark.parametrize("threshold", ["0", "100000"])
def test_collision_estimator_and_distribcell_mg(emu_lib, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    _check_generation(lib.Engine(emu_lib), c5g7_with_estimators, 1500, 1e-11)


def _analog_case(name):
    """the analog-estimator cases of tests/reference_runs.py (on which the oracle equals the
    reference's own binary bit for bit): analog material x energy and mesh tallies with all six
    scores + a collision-estimator tally, CE / MG, with and without survival biasing"""
    import reference_runs as rr
    m, libx, model, box, of, ce = rr.build(name)
    return model, box, of


ANALOG = ["pincell_ce_analog", "pincell_ce_analog_survival", "pincell_mg_analog", "pincell_mg_analog_survival"]


@pytest.mark.parametrize("sweep", ["1", "2"])
@pytest.mark.parametrize("threshold", ["0", "100000"])
@pytest.mark.parametrize("name", ANALOG)
def test_analog_estimator(emu_lib, name, threshold, sweep, monkeypatch):
    """score_analog_tally_ce / _mg (src/tallies/tally_scoring.cpp:2314-2414): scored after the
    collision from the collision records (list kernels: k_after_collision_score; tail: in line)"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)
    _check_generation(lib.Engine(emu_lib), lambda: _analog_case(name), 1500, 1e-11)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (container only)")
@pytest.mark.parametrize("name", ["assembly", "c5g7", "pwr_core"])
def test_distribcell_offsets_equal_reference(name):
    """cell instances of every material cell with the reference's own offset tables
    (prepare_distribcell with material_cell_offsets, then Cell instance of src/geometry.cpp:66-98)
    against the oracle walking the flattener's tables"""
    import cases
    import ref_geometry
    builders = {"assembly": lambda: examples.assembly_ce(n_grid=300, tallies=False),
                "c5g7": lambda: examples.c5g7_core(mesh_dim=17, tallies=False),
                "pwr_core": lambda: examples.pwr_core_ce(n_grid=300, n_fuel_nuclides=8, tallies=False)}
    m, libx, box, of = builders[name]()
    m.settings["material_cell_offsets"] = 1
    model = mgxs.build_mg_model(m, libx) if name == "c5g7" else nuclide.build_ce_model(m, libx)
    assert int(model["cell_distribcell_index"].max()) >= 0
    ref = ref_geometry.ReferenceGeometry(m)
    assert ref.lib.ref_geom_prepare_distribcell() == 0, ref.lib.ref_geom_error()
    xyz, uvw = cases.geometry_probe_points(box, 30000)
    o = Oracle(model)
    cell_o, inst_o, mat_o, nc_o = o.find_cell(xyz, uvw)
    cid, inst, mat, nco = ref.find_cell(xyz, uvw)
    assert np.array_equal(inst, inst_o)
    assert inst_o.max() > 100                       # lattices of pins: hundreds of instances


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", ["0", "2048"])
def test_collision_estimator_and_distribcell_on_gpu(gpu_engine, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    _check_generation(gpu_engine, lambda: assembly_with_estimators(4000), 20000, None, exact_bank=False)
    _check_generation(gpu_engine, c5g7_with_estimators, 20000, None, exact_bank=False)


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", ["0", "2048"])
@pytest.mark.parametrize("name", ANALOG)
def test_analog_estimator_on_gpu(gpu_engine, name, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    mg = "_mg_" in name
    # multigroup: the same histories as the oracle (every bin to 1e-9); CE: a few flipped decisions
    _check_generation(gpu_engine, lambda: _analog_case(name), 20000, 1e-9 if mg else None, exact_bank=mg)


@pytest.mark.parametrize("threshold", ["0", "100000"])
@pytest.mark.parametrize("name", ["pincell_ce_density", "assembly_ce_density"])
def test_cell_density_multipliers(emu_lib, name, threshold, monkeypatch):
    """Cell::density_mult (src/cell.cpp:104-111) / Material::atom_density(i, rho): resolved at
    omcgpu_init into (material, multiplier) device materials -- one density per cell (with several
    temperatures per nuclide on top) and one per cell INSTANCE of a lattice; on both the oracle
    equals the reference's binary bit for bit"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    _check_generation(lib.Engine(emu_lib), lambda: _analog_case(name), 1500, 1e-11)


@pytest.mark.parametrize("threshold", ["0", "100000"])
def test_cell_density_multipliers_multigroup(emu_lib, threshold, monkeypatch):
    """Mgxs::calculate_xs scales total / absorption / nu-fission by the multiplier of the cell
    instance (src/mgxs.cpp:620-625): read at the lookup by the extended build (one density per
    instance of the UO2 pins' fuel cell of the C5G7-style core; the oracle equals the reference's
    binary bit for bit on this model)"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model = _check_generation(lib.Engine(emu_lib), lambda: _analog_case("c5g7_core_density"), 1500, 1e-11)
    eng = lib.Engine(emu_lib)
    assert eng.lib.omcgpu_needs_extended_build(model.byref()) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pincell_ce_density", "assembly_ce_density", "c5g7_core_density"])
def test_cell_density_multipliers_on_gpu(gpu_engine, name, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "2048")
    _check_generation(gpu_engine, lambda: _analog_case(name), 20000, None, exact_bank=False)


@pytest.mark.parametrize("sweep", ["1", "2"])
@pytest.mark.parametrize("threshold", ["0", "100000"])
def test_secondary_neutrons_from_integer_yields(emu_lib, threshold, sweep, monkeypatch):
    """(n,2n) in every nuclide of the pin cell: inelastic_scatter creates yield - 1 secondary
    neutrons (src/physics.cpp:1158-1162), event_revive_from_secondary transports them after the
    particle's death (src/particle.cpp:417-473) -- pool of sites + one list per particle on the
    device; the oracle equals the reference's binary bit for bit on this model"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)
    _check_generation(lib.Engine(emu_lib), lambda: _analog_case("pincell_ce_n2n"), 1500, 1e-11)


def test_secondaries_change_the_answer_and_a_full_pool_is_an_error(emu_lib, monkeypatch):
    """the same generation with the pool squeezed to 2 sites fails loudly (nothing is dropped
    silently); and the (n,2n) reactions do happen in this model: hundreds of extra events"""
    import reference_runs as rr
    m, libx, model, box, of, ce = rr.build("pincell_ce_n2n")
    o = Oracle(model)
    n = 1500
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    monkeypatch.setenv("OMCGPU_SECONDARY_POOL", "2")
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, False)
    with pytest.raises(lib.OmcGpuError, match="secondary-neutron pool is full"):
        eng.transport_generation(gi, src)
    eng.finalize()


@pytest.mark.gpu
def test_secondary_neutrons_on_gpu(gpu_engine, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "2048")
    _check_generation(gpu_engine, lambda: _analog_case("pincell_ce_n2n"), 20000, None, exact_bank=False)
