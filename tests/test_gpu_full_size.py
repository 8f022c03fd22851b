"""GPU tier, BASELINE.json's full size (config C2: 10 nuclides x 100 000 points, 10^6
histories per generation, the bench tallies): the oracle needs minutes for one such
generation, so parity is checked through size-independent properties of the path --
sortedness of the fission bank (src/bank.cpp:76-134), event balance, two invariants that
tie the tallies to the global k tallies, run-to-run idempotence, and equality of the
scheduling variants (they must bank the very same sites)."""
import os

import numpy as np
import pytest

from openmc_b200 import examples, lib, nuclide, source

pytestmark = pytest.mark.gpu

N = 1_000_000


@pytest.fixture(scope="module")
def full(gpu_engine):
    m, libx, box, of = examples.pincell_ce(n_grid=100_000)
    model = nuclide.build_ce_model(m, libx)
    gpu_engine.init(model, 0, N)
    src = source.sample_initial_source(gpu_engine, N, box, of)
    gpu_engine.finalize()
    return model, src


def _generation(eng, model, src, env, monkeypatch, tallies=True):
    for k in ("OMCGPU_SWEEP", "OMCGPU_ORDERED", "OMCGPU_TAIL_THRESHOLD"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    eng.init(model, 0, N)
    gi = eng.gen_in(N, 0, 0, 1.0, tallies)
    fb, out = eng.transport_generation(gi, src)
    tal = []
    for t in range(int(model["n_tallies"])):
        n = int(model["tally_result_ofs"][t + 1] - model["tally_result_ofs"][t])
        v = np.zeros(n)
        eng.tally_results(t, v)
        tal.append(v)
    eng.finalize()
    return fb.copy(), out, tal


def test_full_size_properties(gpu_engine, full, monkeypatch):
    model, src = full
    fb, out, tal = _generation(gpu_engine, model, src, {}, monkeypatch)
    assert out.n_lost == 0 and out.n_fission == fb.size and 0.5 * N < fb.size < 3 * N
    # sort_fission_bank: (parent_id, progeny_id) strictly ascending, progeny ids 0..k-1 per parent
    key = fb["parent_id"].astype(np.int64) * 64 + fb["progeny_id"].astype(np.int64)
    assert np.all(np.diff(key) > 0)
    first = np.r_[True, np.diff(fb["parent_id"]) != 0]
    assert np.all(fb["progeny_id"][first] == 0)
    assert np.all((np.diff(fb["progeny_id"]) == 1) | first[1:])
    assert fb["parent_id"].min() >= 1 and fb["parent_id"].max() <= N
    # every lookup is followed by exactly one flight, every flight by one collision or crossing
    assert out.n_xs_lookups == out.n_collisions + out.n_crossings
    assert out.n_events == out.n_collisions + out.n_crossings
    assert out.total_weight == float(N)
    # tracklength estimate of k = sum over the mesh of the nu-fission score (the mesh covers
    # the pin cell); flux summed over the mesh = flux summed over the energy spectrum
    mesh = tal[0].reshape(-1, 3)             # [bin][flux, fission, nu-fission]
    np.testing.assert_allclose(mesh[:, 2].sum(), out.global_tally[2], rtol=1e-9)
    np.testing.assert_allclose(mesh[:, 0].sum(), tal[1].sum(), rtol=1e-6)
    assert np.all(mesh >= 0) and np.count_nonzero(mesh[:, 0]) == mesh.shape[0]
    assert np.all(mesh[:, 1] <= mesh[:, 2])  # nu >= 1


def test_full_size_idempotent_and_schedule_independent(gpu_engine, full, monkeypatch):
    """the same generation again, and under every scheduling variant (merged alive list,
    paired lists with appended / index-ordered lists, all-tail): identical banked sites and
    k tallies bit for bit -- scheduling must not touch a history's arithmetic"""
    model, src = full
    ref_fb, ref_out, ref_tal = _generation(gpu_engine, model, src, {}, monkeypatch)
    variants = [{}, {"OMCGPU_SWEEP": "1"}, {"OMCGPU_ORDERED": "0"}, {"OMCGPU_TAIL_THRESHOLD": "200000"}]
    for env in variants:
        fb, out, tal = _generation(gpu_engine, model, src, env, monkeypatch)
        assert fb.tobytes() == ref_fb.tobytes(), env
        assert out.n_events == ref_out.n_events and out.n_collisions == ref_out.n_collisions, env
        # per-history accumulators are summed in a fixed order (k_reduce): bit-identical
        assert list(out.global_tally) == list(ref_out.global_tally), env
        # tallies: RED order differs from run to run -> equal to rounding
        for a, b in zip(tal, ref_tal):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_full_size_next_generation_equals_host_comb(gpu_engine, full, monkeypatch):
    """omcgpu_next_generation (device sort + comb, page-locked bank) leaves in the host bank
    the sites synchronize_bank's comb (src/eigenvalue.cpp:152-178) selects on the host"""
    model, src = full
    fb, out, _ = _generation(gpu_engine, model, src, {}, monkeypatch, tallies=False)
    want = np.ascontiguousarray(fb[lib.comb_indices(0, fb.size, fb.size, N, 1)])
    gpu_engine.init(model, 0, N)
    bank = src.copy()
    gpu_engine.host_register(bank)
    out2 = gpu_engine.next_generation(gpu_engine.gen_in(N, 0, 0, 1.0, False), bank, 1)
    gpu_engine.host_unregister(bank)
    gpu_engine.finalize()
    assert out2.n_fission == out.n_fission
    assert bank.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------
# BASELINE.json's full sizes BY VALUE.  Histories are independent (ids and seeds are
# global, src/simulation.cpp:592,613-616), so the oracle transports a 20 000-history
# slice of the very same 10^6-history generation with id_offset = the slice's start,
# and the sites those parents banked must be the ones the GPU banked for them.
# ------------------------------------------------------------------------------
def test_full_size_generation_by_value(gpu_engine, full, monkeypatch):
    """config C2 at BASELINE size (100 000-point grids, 10^6 histories, tallies active): the
    fission sites of a 20 000-history slice equal the oracle's, parent by parent"""
    from oracle.binding import Oracle
    model, src = full
    fb, out, _ = _generation(gpu_engine, model, src, {}, monkeypatch)
    o = Oracle(model)
    lo, n = 437_000, 20_000
    gi = lib.Engine.gen_in(None, n, lo, 0, 1.0, True)
    fb_o, out_o = o.transport_generation(gi, np.ascontiguousarray(src[lo:lo + n]))
    assert out_o.n_lost == 0
    import cases
    same, close = cases.bank_parity(fb, fb_o, lo, lo + n)
    print(f"parents with identical site counts {same:.5f}, sites equal {close:.5f}")


def test_full_size_many_nuclide_lookup_by_value(gpu_engine):
    """config C5 at BASELINE size (300-nuclide fuel x 100 000 points, 1.8 GB of tables):
    grid indices bit-exact and macroscopic cross sections <= 1e-12 for every material"""
    import cases
    from oracle.binding import Oracle
    m, libx, box, of = examples.pwr_core_ce(n_grid=100_000, n_fuel_nuclides=300)
    model = nuclide.build_ce_model(m, libx)
    o = Oracle(model)
    n_nuc = np.diff(np.asarray(model["mat_nuc_ofs"]))
    gpu_engine.init(model, 0, 1024)
    E = cases.xs_probe_energies(model, None, 20000)
    for mat in (int(np.argmax(n_nuc)), int(np.argmin(n_nuc)), int(model["n_materials"]) - 1):
        mats = np.full(E.size, mat, np.int32)
        xs_g, ig_g = gpu_engine.calculate_xs(E, mats, int(n_nuc[mat]))
        xs_o, ig_o = o.calculate_xs(E, mats, int(n_nuc[mat]))
        assert np.array_equal(ig_g, ig_o)
        np.testing.assert_allclose(xs_g, xs_o, rtol=1e-12, atol=0)
    gpu_engine.finalize()
