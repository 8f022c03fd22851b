"""GPU tier: the CUDA path (libomcgpu.so through the C ABI) against the oracle.

Bars (BASELINE.json north_star): integer / index work bit-exact; cross sections
<= 1e-12 relative; k-eff and tallies statistically consistent (here, for the MG
goldens, they reproduce the reference's printed digits)."""
import numpy as np
import pytest

import cases
from openmc_b200 import lib
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu


def test_rng_bit_exact(gpu_engine):
    model, _, _ = cases.build("mg_legendre")
    o = Oracle(model)
    g = np.random.default_rng(3)
    seeds = g.integers(0, 2**63, 100000, dtype=np.uint64)
    ns = g.integers(0, 2**63, 100000, dtype=np.uint64)
    assert np.array_equal(gpu_engine.future_seed(ns, seeds), o.future_seed(ns, seeds))
    assert np.array_equal(gpu_engine.prn(seeds, 8), o.prn(seeds, 8))   # bit-exact doubles


@pytest.mark.parametrize("name", ["mg_legendre", "mg_temperature_multi"])
def test_geometry_bit_exact(gpu_engine, name):
    model, box, _ = cases.build(name)
    o = Oracle(model)
    gpu_engine.init(model, 0, 200000)
    g = np.random.default_rng(4)
    n = 200000
    box = np.asarray(box, float)
    lo, hi = box[:3] - 0.05 * (box[3:] - box[:3]), box[3:] + 0.05 * (box[3:] - box[:3])
    xyz = lo + g.uniform(size=(n, 3)) * (hi - lo)
    uvw = g.normal(size=(n, 3)); uvw /= np.linalg.norm(uvw, axis=1)[:, None]
    for a, b in zip(gpu_engine.find_cell(xyz, uvw), o.find_cell(xyz, uvw)):
        assert np.array_equal(a, b)
    dg, sg, lg, tg = gpu_engine.distance_to_boundary(xyz, uvw)
    do, so, lo_, to = o.distance_to_boundary(xyz, uvw)
    assert np.array_equal(sg, so) and np.array_equal(lg, lo_) and np.array_equal(tg, to)
    np.testing.assert_allclose(dg, do, rtol=1e-12)
    gpu_engine.finalize()


@pytest.mark.parametrize("threshold,sweep", [("0", "2"), ("2048", "2"), ("0", "1")])
@pytest.mark.parametrize("name", list(cases.GOLDEN))
def test_reference_goldens_on_gpu(gpu_engine, name, threshold, sweep, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    monkeypatch.setenv("OMCGPU_SWEEP", sweep)   # paired lists (default) | merged alive list
    model, box, of = cases.build(name)
    src = Oracle(model).sample_initial_source(1000, box, of)
    sim = lib.Simulation(gpu_engine, model, 1000, 10, 5, src)
    k = sim.run()
    kg, sg, tallies = cases.GOLDEN[name]
    assert sim.stats["n_lost"] == 0
    assert abs(k[0] - float(kg)) < 5e-7 * float(kg) and abs(k[1] - float(sg)) < 5e-6 * float(sg)
    for t, (ts, tsq) in enumerate(tallies):
        assert abs(sim.tally_results[t][:, :, 1].sum() - float(ts)) < 1e-6 * float(ts)
        assert abs(sim.tally_results[t][:, :, 2].sum() - float(tsq)) < 2e-6 * float(tsq)
    sim.simulation_finalize()


def test_generation_matches_oracle(gpu_engine, monkeypatch):
    """One generation through host buffers (omcgpu_transport_generation): the sorted
    fission bank has the same sites as the oracle's (integer fields exact, reals to
    1e-12), the k tallies agree to 1e-12 relative."""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "0")
    model, box, of = cases.build("mg_temperature_multi")
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_fission == out_o.n_fission
    assert out.n_collisions == out_o.n_collisions and out.n_crossings == out_o.n_crossings
    for f in ("delayed_group", "parent_id", "progeny_id", "E"):
        assert np.array_equal(fb[f], fb_o[f]), f
    np.testing.assert_allclose(fb["r"], fb_o["r"], rtol=1e-9, atol=1e-9)   # ulp drift over many flights
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-12)
    for t in range(2):
        m = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(m); gpu_engine.tally_results(t, v)
        np.testing.assert_allclose(v, o.tally_results(t, m), rtol=1e-11)
    gpu_engine.finalize()


@pytest.mark.parametrize("threshold", ["0", "2048"])
def test_pincell_multigroup_generation_matches_oracle(gpu_engine, threshold, monkeypatch):
    """config C1 (examples/pincell_multigroup: 7-group C5G7 data, reflective pin cell, 100 x 100
    mesh {flux, fission, nu-fission} + 500-bin spectrum tallies): one generation against the oracle"""
    from openmc_b200 import examples, mgxs
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    m, libx, box, of = examples.pincell_multigroup()
    model = mgxs.build_mg_model(m, libx)
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    assert abs(out.n_fission - out_o.n_fission) <= 5
    assert abs(out.n_events - out_o.n_events) <= 1e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-4)
    if out.n_fission == out_o.n_fission:
        assert np.mean(fb["parent_id"] == fb_o["parent_id"]) > 0.99
    for t in range(int(model["n_tallies"])):
        k = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(k); gpu_engine.tally_results(t, v)
        vo = o.tally_results(t, k)
        assert abs(v.sum() - vo.sum()) < 1e-4 * vo.sum()
    gpu_engine.finalize()


def test_second_scoring_kernel_equals_the_first(gpu_engine, monkeypatch):
    """k_score2 (OMCGPU_SCORE_VERSION=2: ray constants once per flight, short parts started by a
    multiply, lanes in step) scores the same bins with the same lengths as k_score: every bin of
    the C1 tallies to 1e-10 relative (the two differ by the rounding of the crossing distances)."""
    from openmc_b200 import examples, mgxs
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "2048")
    m, libx, box, of = examples.pincell_multigroup()
    model = mgxs.build_mg_model(m, libx)
    n = 20000
    src = Oracle(model).sample_initial_source(n, box, of)
    res = {}
    for version in ("1", "2"):
        monkeypatch.setenv("OMCGPU_SCORE_VERSION", version)
        gpu_engine.init(model, 0, n)
        gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
        fb, out = gpu_engine.transport_generation(gi, src)
        tall = []
        for t in range(int(model["n_tallies"])):
            k = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
            v = np.zeros(k); gpu_engine.tally_results(t, v)
            tall.append(v)
        res[version] = (out.n_fission, out.n_events, tall)
        gpu_engine.finalize()
    assert res["1"][:2] == res["2"][:2]
    for a, b in zip(res["1"][2], res["2"][2]):
        assert a.sum() > 0
        np.testing.assert_allclose(b, a, rtol=1e-10, atol=1e-12 * a.max())


def test_fission_sites_counted_on_a_mesh_on_gpu(gpu_engine):
    """omcgpu_count_fission_sites (Mesh::count_sites for the Shannon entropy of the drop-in): the
    device's counts equal counting the fission bank the host reads back"""
    model, box, of = cases.build("mg_temperature_multi")
    n = 20000
    src = Oracle(model).sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, False)
    fb, out = gpu_engine.transport_generation(gi, src)
    r = fb["r"]
    lo, hi = r.min(axis=0) - 1e-9, r.max(axis=0) + 1e-9
    shape = (7, 5, 3)
    width = (hi - lo) / np.array(shape)
    counts, outside = gpu_engine.count_fission_sites(shape, lo, width)
    idx = np.ceil((r - lo) / width).astype(int) - 1
    want = np.zeros(int(np.prod(shape)))
    np.add.at(want, idx[:, 0] + shape[0] * (idx[:, 1] + shape[1] * idx[:, 2]), fb["wgt"])
    assert not outside and np.array_equal(counts, want) and counts.sum() == out.n_fission
    gpu_engine.finalize()


def test_device_collision_equals_reference_vectors_on_gpu(gpu_engine):
    """collision_ce ON THE DEVICE (omcgpu_collide) against tests/golden/refcoll_ce.npz, the
    vectors the reference's collision() wrote.  Integer results (seed = the number of random
    numbers consumed, fission-site counts, delayed groups) are identical; reals agree to the
    last-ulp differences between the device libm and glibc (1e-12 relative)."""
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    model = cases.load_model_npz(os.path.join(golden, "ce_model.npz"))
    v = np.load(os.path.join(golden, "refcoll_ce.npz"))
    mat, E, u, seeds = (np.ascontiguousarray(v[k]) for k in ("material", "E_in", "u_in", "seed_in"))
    n = E.size
    for tag, survival in (("a", 0), ("s", 1)):
        model.struct.survival_biasing = survival
        gpu_engine.init(model, 0, max(n, 1024))
        r = gpu_engine.collide(mat, E, u, seeds, float(v["k_eff"]), float(v[f"{tag}_wgt0"]))
        gpu_engine.finalize()
        ofs = np.concatenate([[0], np.cumsum(r["n_sites"])]).astype(np.int64)
        # a last-ulp libm difference can flip an accept / reject decision of a rejection loop:
        # such a neutron draws a different number of random numbers.  They must be rare.
        same_seed = r["seed"] == v[f"{tag}_seed"]
        assert same_seed.mean() > 0.999, same_seed.mean()
        ok = same_seed
        np.testing.assert_allclose(r["E"][ok], v[f"{tag}_E"][ok], rtol=1e-12, atol=0)
        np.testing.assert_allclose(r["u"][ok], v[f"{tag}_u"][ok], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["wgt"][ok], v[f"{tag}_wgt"][ok], rtol=1e-12, atol=0)
        np.testing.assert_allclose(r["k_abs"][ok], v[f"{tag}_k_abs"][ok], rtol=1e-12, atol=0)
        n_ref = np.diff(v[f"{tag}_site_ofs"])
        assert np.array_equal(r["n_sites"][ok], n_ref[ok])
        if same_seed.all():
            assert np.array_equal(ofs, v[f"{tag}_site_ofs"])
            assert np.array_equal(r["sites"]["delayed_group"], v[f"{tag}_site_group"])
            np.testing.assert_allclose(r["sites"]["E"], v[f"{tag}_site_E"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(r["sites"]["u"], v[f"{tag}_site_u"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(r["sites"]["time"], v[f"{tag}_site_time"], rtol=1e-12, atol=1e-300)


# ------------------------------------------------------------------------------
# continuous energy (config C2 geometry, synthetic library)
# ------------------------------------------------------------------------------
def test_ce_macro_xs_and_grid_index(gpu_engine):
    """energy-grid indices bit-exact, macroscopic cross sections <= 1e-12 relative
    (north_star), including energies on grid points and on hash-bin edges."""
    model, box, of, libx = cases.build_ce(n_grid=20000)
    o = Oracle(model)
    gpu_engine.init(model, 0, 1024)
    E = cases.xs_probe_energies(model, libx, 400000)
    for mat in range(int(model["n_materials"])):
        mats = np.full(E.size, mat, np.int32)
        xs_g, ig_g = gpu_engine.calculate_xs(E, mats, 4)
        xs_o, ig_o = o.calculate_xs(E, mats, 4)
        assert np.array_equal(ig_g, ig_o)
        np.testing.assert_allclose(xs_g, xs_o, rtol=1e-12, atol=0)
    gpu_engine.finalize()


@pytest.mark.parametrize("threshold", ["0", "2048"])
def test_ce_generation_matches_oracle(gpu_engine, threshold, monkeypatch):
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model, box, of, _ = cases.build_ce(n_grid=20000)
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    # libm last-ulp differences may flip a handful of discrete decisions in ~2e6
    # events; the banks must agree except for those few histories
    assert abs(out.n_fission - out_o.n_fission) <= 20
    assert abs(out.n_events - out_o.n_events) <= 2e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=2e-4)
    cases.bank_parity(fb, fb_o, 0, n)      # parent by parent: same sites but for the few flipped histories
    gpu_engine.finalize()


def test_ce_eigenvalue_statistics(gpu_engine):
    """k-eff within 3 combined standard deviations of the oracle's; tally means pass
    a z-test (north_star) -- independent seeds so the comparison is statistical."""
    model, box, of, _ = cases.build_ce(n_grid=20000, mesh_dim=4, energy_bins=8)
    o = Oracle(model)
    n, nb, ni = 4000, 14, 4
    src = o.sample_initial_source(n, box, of)
    r = o.run_eigenvalue(src, nb, ni, master_seed=1)
    sim = lib.Simulation(gpu_engine, model, n, nb, ni, src, seed=12345)
    k = sim.run()
    sig = np.hypot(k[1], r["k_combined"][1])
    assert abs(k[0] - r["k_combined"][0]) < 3 * sig, (k, r["k_combined"])
    nreal = nb - ni
    for t in range(2):
        mean_g, std_g = sim.tally_mean_std(t)
        lo, hi = int(model["tally_result_ofs"][t]), int(model["tally_result_ofs"][t + 1])
        mean_o = r["tally_sum"][lo:hi] / nreal
        var_o = np.maximum(r["tally_sum_sq"][lo:hi] / nreal - mean_o ** 2, 0) / (nreal - 1)
        z = (mean_g.ravel() - mean_o) / np.sqrt(std_g.ravel() ** 2 + var_o + 1e-300)
        ok = np.isfinite(z) & (mean_o > 0)
        assert np.all(np.abs(z[ok]) < 5.0), np.abs(z[ok]).max()
        assert np.mean(z[ok] ** 2) < 2.5
    sim.simulation_finalize()


# ------------------------------------------------------------------------------
# lattices (configs C3 / C4)
# ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["assembly", "c5g7", "pwr_core"])
def test_lattice_geometry_bit_exact(gpu_engine, name):
    """cell / instance / material / level, crossing surface tokens and lattice
    translations bit-exact on 2- and 3-level lattice geometry, incl. points placed
    on lattice planes; distances to 1e-12"""
    model, box, of = cases.build_lattice(name)
    o = Oracle(model)
    gpu_engine.init(model, 0, 200000)
    xyz, uvw = cases.geometry_probe_points(box, 200000)
    for a, b in zip(gpu_engine.find_cell(xyz, uvw), o.find_cell(xyz, uvw)):
        assert np.array_equal(a, b)
    dg, sg, lg, tg = gpu_engine.distance_to_boundary(xyz, uvw)
    do, so, lo_, to = o.distance_to_boundary(xyz, uvw)
    assert np.array_equal(sg, so) and np.array_equal(lg, lo_) and np.array_equal(tg, to)
    np.testing.assert_allclose(dg, do, rtol=1e-12)
    gpu_engine.finalize()


def test_many_nuclide_lookup(gpu_engine):
    """config C5 fuel (40 nuclides here) through the lockstep lookup kernels: grid indices
    bit-exact, macroscopic cross sections to 1e-12 (north_star)"""
    model, box, of = cases.build_lattice("pwr_core")
    n_nuc = np.diff(np.asarray(model["mat_nuc_ofs"]))
    mat = int(np.argmax(n_nuc))
    o = Oracle(model)
    gpu_engine.init(model, 0, 1024)
    E = cases.xs_probe_energies(model, None, 100000)
    mats = np.full(E.size, mat, np.int32)
    xs_g, ig_g = gpu_engine.calculate_xs(E, mats, int(n_nuc[mat]))
    xs_o, ig_o = o.calculate_xs(E, mats, int(n_nuc[mat]))
    assert np.array_equal(ig_g, ig_o)
    np.testing.assert_allclose(xs_g, xs_o, rtol=1e-12, atol=0)
    gpu_engine.finalize()


@pytest.mark.parametrize("threshold", ["0", "4096"])
def test_pwr_core_generation_matches_oracle(gpu_engine, threshold, monkeypatch):
    """config C5 (full core, 3 levels, 40-nuclide fuel): one generation against the oracle"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model, box, of = cases.build_lattice("pwr_core")
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    assert abs(out.n_fission - out_o.n_fission) <= 20
    assert abs(out.n_events - out_o.n_events) <= 2e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=2e-4)
    cases.bank_parity(fb, fb_o, 0, n)
    gpu_engine.finalize()


@pytest.mark.parametrize("threshold", ["0", "16384"])
def test_c5g7_generation_matches_oracle(gpu_engine, threshold, monkeypatch):
    """lattice-of-lattices MG core: same sorted fission bank as the oracle"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", threshold)
    model, box, of = cases.build_lattice("c5g7")
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, of)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, True)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    # a last-ulp libm difference may flip a rare discrete decision
    assert abs(out.n_fission - out_o.n_fission) <= 5
    assert abs(out.n_events - out_o.n_events) <= 1e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=1e-4)
    if out.n_fission == out_o.n_fission:
        same = np.mean(fb["parent_id"] == fb_o["parent_id"])
        assert same > 0.99
    m = int(model["tally_result_ofs"][1])
    v = np.zeros(m); gpu_engine.tally_results(0, v)
    vo = o.tally_results(0, m)
    assert abs(v.sum() - vo.sum()) < 1e-4 * vo.sum()
    gpu_engine.finalize()


def test_assembly_eigenvalue_statistics(gpu_engine):
    """config C3 (17 x 17 CE assembly): k within 3 combined sigma of the oracle,
    pin-mesh tallies pass a z-test"""
    model, box, of = cases.build_lattice("assembly")
    o = Oracle(model)
    n, nb, ni = 4000, 12, 4
    src = o.sample_initial_source(n, box, of)
    r = o.run_eigenvalue(src, nb, ni, master_seed=1)
    sim = lib.Simulation(gpu_engine, model, n, nb, ni, src, seed=777)
    k = sim.run()
    assert sim.stats["n_lost"] == 0
    sig = np.hypot(k[1], r["k_combined"][1])
    assert abs(k[0] - r["k_combined"][0]) < 3 * sig, (k, r["k_combined"])
    nreal = nb - ni
    mean_g, std_g = sim.tally_mean_std(0)
    hi = int(model["tally_result_ofs"][1])
    mean_o = r["tally_sum"][:hi] / nreal
    var_o = np.maximum(r["tally_sum_sq"][:hi] / nreal - mean_o ** 2, 0) / (nreal - 1)
    z = (mean_g.ravel() - mean_o) / np.sqrt(std_g.ravel() ** 2 + var_o + 1e-300)
    ok = np.isfinite(z) & (mean_o > 0)
    assert np.all(np.abs(z[ok]) < 5.0), np.abs(z[ok]).max()
    assert np.mean(z[ok] ** 2) < 2.5
    sim.simulation_finalize()


@pytest.mark.parametrize("case", ["surface_zoo", "rotated_fills"])
def test_surface_zoo_generation_matches_oracle(gpu_engine, case, monkeypatch):
    """spheres, cones, x / y cylinders, a general plane, union / complement regions, vacuum
    sphere; translated + rotated universe and lattice fills: one generation against the oracle"""
    import reference_cases as rc
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "2048")
    m, model, box = rc.build(case)
    o = Oracle(model)
    n = 20000
    src = o.sample_initial_source(n, box, False)
    gpu_engine.init(model, 0, n)
    gi = gpu_engine.gen_in(n, 0, 0, 1.0, False)
    fb, out = gpu_engine.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src)
    assert out.n_lost == 0
    assert abs(out.n_fission - out_o.n_fission) <= 5
    assert abs(out.n_events - out_o.n_events) <= 2e-4 * out_o.n_events
    np.testing.assert_allclose(out.global_tally[:], out_o.global_tally[:], rtol=2e-4)
    gpu_engine.finalize()
