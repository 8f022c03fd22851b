"""CPU tier: the engine sources (openmc_b200/csrc), host-compiled for testing
only, driven through the same C ABI and the same Python host driver as the GPU
library.  Checks the event loop / work lists / bank sort / comb / host driver
against the reference goldens and against the oracle, without a GPU."""
import ctypes as C

import numpy as np
import pytest

import cases
from openmc_b200 import _abi as abi
from openmc_b200 import lib
from oracle.binding import Oracle


@pytest.mark.parametrize("sweep", ["1", "2"])   # merged alive list | paired collision / crossing lists
@pytest.mark.parametrize("threshold", ["0", "64", "100000"])
@pytest.mark.parametrize("name", ["mg_legendre", "mg_survival_biasing", "mg_temperature_multi", "mg_basic"])
def test_golden_through_engine(emu_lib, name, threshold, sweep, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)
    model, box, of = cases.build(name)
    src = Oracle(model).sample_initial_source(1000, box, of)
    sim = lib.Simulation(lib.Engine(emu_lib), model, 1000, 10, 5, src)
    k = sim.run()
    kg, sg, tallies = cases.GOLDEN[name]
    assert cases.fmt(k[0]) == kg and cases.fmt(k[1]) == sg
    for t, (ts, tsq) in enumerate(tallies):
        assert cases.fmt(sim.tally_results[t][:, :, 1].sum()) == ts
        assert cases.fmt(sim.tally_results[t][:, :, 2].sum()) == tsq
    sim.simulation_finalize()


@pytest.mark.parametrize("score_version", [1, 2])
def test_host_buffer_generation_matches_oracle(emu_lib, monkeypatch, score_version):
    """omcgpu_transport_generation == oracle generation: identical sorted fission bank.
    score_version 2: the arithmetic of k_score2 (ray constants per flight, parts started by a multiply)."""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "0")
    monkeypatch.setenv("OMCGPU_SCORE_VERSION", str(score_version))
    model, box, of = cases.build("mg_temperature_multi")
    o = Oracle(model)
    src = o.sample_initial_source(2000, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 4096)
    gi = eng.gen_in(2000, 0, 0, 1.0, True)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    assert out.n_collisions == out_o.n_collisions and out.n_crossings == out_o.n_crossings
    for f in ("r", "u", "E", "time", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-13)
    for t in range(2):
        n = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(n); eng.tally_results(t, v)
        np.testing.assert_allclose(v, o.tally_results(t, n), rtol=1e-12)
    eng.finalize()


def test_resident_equals_host_buffers(emu_lib):
    model, box, of = cases.build("mg_legendre")
    src = Oracle(model).sample_initial_source(500, box, of)
    ks = []
    for resident in (True, False):
        sim = lib.Simulation(lib.Engine(emu_lib), model, 500, 6, 2, src, resident=resident)
        ks.append(sim.run())
        sim.simulation_finalize()
    assert ks[0] == ks[1]


def test_next_generation_equals_transport_plus_host_comb(emu_lib):
    """omcgpu_next_generation (device comb, source bank replaced in place) selects the
    very sites that transport_generation + the host-side comb of synchronize_bank do."""
    model, box, of = cases.build("mg_legendre")
    src = Oracle(model).sample_initial_source(1500, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 2048)
    a, b = src.copy(), src.copy()
    eng.host_register(b)
    for g in range(3):
        gi = eng.gen_in(1500, 0, g * 1500, 1.0, False)
        fb, out_a = eng.transport_generation(gi, a)
        a = np.ascontiguousarray(fb[lib.comb_indices(0, fb.size, fb.size, 1500, g + 1)])
        out_b = eng.next_generation(gi, b, g + 1)
        assert out_a.n_fission == out_b.n_fission
        assert a.tobytes() == b.tobytes()
    eng.host_unregister(b)
    eng.finalize()


def test_comb_teeth_equal_the_recurrence(emu_lib):
    """omcgpu_comb_indices fills the teeth binade by binade (integer progressions); the
    selected sites must equal the reference's recurrence tooth += teeth_distance exactly --
    also when teeth_distance has few mantissa bits (ties), is < 1 or > 2, on rank slices"""
    eng = lib.Engine(emu_lib)
    g = np.random.default_rng(3)
    cases_ = [(1234, 1000), (999, 1000), (3000, 1000), (1000, 1000), (1250, 1000), (1500, 1000),
              (2_000_000, 1 << 20), ((1 << 21) + 1, 1 << 20), (1_048_577, 1 << 20), (5, 4000), (4_123_457, 4_000_000)]
    cases_ += [(int(t), int(n)) for t, n in zip(g.integers(1, 3_000_000, 12), g.integers(1, 1_500_000, 12))]
    for total, n in cases_:
        for seed_id in (1, 7, 123):
            a = lib.comb_indices(0, total, total, n, seed_id)
            b = eng.comb_indices(0, total, total, n, seed_id)
            assert np.array_equal(a, b), (total, n, seed_id)
        lo = total // 3                       # a rank that holds sites [lo, lo + n_local)
        n_local = total // 2
        a = lib.comb_indices(lo, n_local, total, n, 5)
        b = eng.comb_indices(lo, n_local, total, n, 5)
        assert np.array_equal(a, b), (total, n, "slice")


def test_comb_matches_oracle(emu_lib):
    model, box, of = cases.build("mg_legendre")
    o = Oracle(model)
    for total, n in ((1234, 1000), (999, 1000), (3000, 1000), (1000, 1000)):
        a = o.comb_indices(0, total, total, n, seed_id=7)
        b = lib.comb_indices(0, total, total, n, seed_id=7)
        assert np.array_equal(a, b) and a.size == n


@pytest.mark.parametrize("threshold", ["0", "100000"])
@pytest.mark.parametrize("case", ["surface_zoo", "rotated_fills", "periodic_box"])
def test_surface_zoo_generation_matches_oracle(emu_lib, case, threshold, monkeypatch):
    """one generation through spheres, cones, x / y cylinders, a general plane, union and
    complement regions and a vacuum sphere -- and through translated + rotated universe and
    lattice fills (directions re-rotated after every collision, src/particle.cpp:396-406), and
    through translational periodic boundaries:
    same events, same sorted fission bank"""
    import reference_cases as rc
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    m, model, box = rc.build(case)
    o = Oracle(model)
    n = 1500
    src = o.sample_initial_source(n, box, False)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 2048)
    gi = eng.gen_in(n, 0, 0, 1.0, False)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0 and out_o.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    assert out.n_crossings == out_o.n_crossings and out.n_collisions == out_o.n_collisions
    assert (out.global_tally[3] > 0) == (case == "surface_zoo")   # leakage through the vacuum sphere
    for f in ("r", "u", "E", "wgt", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-13)
    eng.finalize()


@pytest.mark.parametrize("sweep", ["1", "2"])
def test_ce_survival_biasing_generation_matches_oracle(emu_lib, sweep, monkeypatch):
    """CE pincell with survival biasing (implicit absorption + russian roulette in collision_ce,
    src/physics.cpp:638-651, 152-167): the engine sources bank the oracle's fission sites and
    score the same tallies -- the oracle's collision is itself held to the reference's
    collision() with survival biasing (test_whole_collision_equals_reference)"""
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)
    model, box, of, libx = cases.build_ce(n_grid=500)
    model.struct.survival_biasing = 1             # (cached model: put it back)
    try:
        o = Oracle(model)
        n = 1500
        src = o.sample_initial_source(n, box, of)
        eng = lib.Engine(emu_lib)
        eng.init(model, 0, 4096)
        gi = eng.gen_in(n, 0, 0, 1.0, True)
        fb, out = eng.transport_generation(gi, src)
        fb_o, out_o = o.transport_generation(gi, src)
    finally:
        model.struct.survival_biasing = 0
    assert out.n_lost == 0 and out_o.n_lost == 0
    assert out.n_fission == out_o.n_fission and out.n_events == out_o.n_events
    assert out.n_crossings == out_o.n_crossings and out.n_collisions == out_o.n_collisions
    for f in ("r", "u", "E", "time", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-12)
    m = int(model["tally_result_ofs"][1])
    v = np.zeros(m); eng.tally_results(0, v)
    np.testing.assert_allclose(v, o.tally_results(0, m), rtol=1e-11, atol=1e-300)
    eng.finalize()


def test_lost_particles_are_counted_and_recorded(emu_lib):
    """Particle::mark_as_lost (src/particle.cpp:762-785): a pin cell whose moderator ends at x = 0.5
    (nothing defined up to the reflective plane at 0.63) loses every particle that crosses that
    plane -- same count as the oracle, and omcgpu_lost_particles returns {id, r, u, E} of the
    first ones at the moment they were lost"""
    from openmc_b200 import examples, mgxs
    m, libx, box, of = examples.pincell_multigroup(mesh_dim=4, energy_bins=4)
    xp = m.surface("x-plane", [0.5]).id
    m.cells[-1].region += f" -{xp}"
    model = mgxs.build_mg_model(m, libx)
    o = Oracle(model)
    n = 400
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, False)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == out_o.n_lost > 10
    rec = eng.lost_particles()
    assert rec.shape == (min(out.n_lost, 64), 8)
    ids = rec[:, 0].astype(np.int64)
    assert len(set(ids)) == len(ids) and ids.min() >= 1 and ids.max() <= n
    # lost on the plane x = 0.5 moving outwards, or leaving the fuel (r = 0.54) beyond it
    # (recorded after the last attempt: TINY_BIT = 1e-8 beyond the surface, src/particle.cpp:686-695)
    on_plane = np.abs(rec[:, 1] - 0.5) < 2e-8
    assert on_plane.sum() > 5 and np.all(rec[on_plane, 4] > 0)
    np.testing.assert_allclose(np.hypot(rec[~on_plane, 1], rec[~on_plane, 2]), 0.54, atol=2e-8)
    assert np.all(rec[:, 1] > 0.5 - 1e-9)
    np.testing.assert_allclose(np.linalg.norm(rec[:, 4:7], axis=1), 1.0, rtol=1e-12)
    assert rec[:3].shape == eng.lost_particles(3).shape
    eng.finalize()


def test_fission_sites_counted_on_a_mesh(emu_lib):
    """Mesh::count_sites (src/mesh.cpp:941-997) over the device fission bank, what shannon_entropy
    (src/eigenvalue.cpp:579-607) is computed from in the drop-in: equal to counting the sites the
    host reads back, sites outside the mesh flagged"""
    model, box, of = cases.build("mg_temperature_multi")
    o = Oracle(model)
    n = 3000
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, False)
    out = eng.run_generation(gi) if False else None
    fb, out = eng.transport_generation(gi, src)
    r = fb["r"]
    lo, hi = r.min(axis=0) - 1e-9, r.max(axis=0) + 1e-9
    for shape in ((4, 3, 2), (5, 5)):
        nd = len(shape)
        width = (hi[:nd] - lo[:nd]) / np.array(shape)
        counts, outside = eng.count_fission_sites(shape, lo[:nd], width)
        idx = np.ceil((r[:, :nd] - lo[:nd]) / width).astype(int)
        want = np.zeros(shape[::-1]).reshape(-1)
        flat = np.zeros(len(r), int)
        cells = 1
        for k in range(nd):
            flat += (idx[:, k] - 1) * cells
            cells *= shape[k]
        np.add.at(want, flat, fb["wgt"])
        assert not outside and np.array_equal(counts, want) and counts.sum() == out.n_fission
    # half the box: the sites beyond it are reported, the others still counted
    width = (hi - lo) / np.array((4, 3, 2))
    width[0] /= 2
    counts, outside = eng.count_fission_sites((4, 3, 2), lo, width)
    assert outside and 0 < counts.sum() < out.n_fission
    eng.finalize()
