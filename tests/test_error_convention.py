"""Error convention of the C ABI (include/omcgpu.h: 0 or a negative OMCGPU_E_* with the text in
omcgpu_last_error(), the way the C API returns OPENMC_E_* with openmc_err_msg,
src/error.cpp:20-33, openmc/lib/error.py:8-44) and host-bank authority between generations
(tests/unit_tests/test_lib.py:451-455 edits the source bank through the C API).  CPU tier:
the engine sources host-compiled for the tests."""
import ctypes as C

import numpy as np
import pytest

import cases
from openmc_b200 import _abi as abi
from openmc_b200 import lib
from oracle.binding import Oracle


def _code(exc):
    return int(str(exc.value).split("]")[0].strip("["))


def test_calls_before_init_are_refused(emu_lib):
    eng = lib.Engine(emu_lib)
    eng.finalize()
    with pytest.raises(lib.OmcGpuError) as e:
        eng.run_generation(eng.gen_in(10, 0, 0, 1.0, False))
    assert _code(e) == abi.OMCGPU_E_UNASSIGNED
    with pytest.raises(lib.OmcGpuError) as e:
        eng.set_source_bank(np.zeros(4, abi.site_dtype))
    assert _code(e) == abi.OMCGPU_E_UNASSIGNED


def test_unsupported_features_fail_at_init(emu_lib):
    """no silent fallback: a surface kind / boundary condition the device path does not
    implement stops omcgpu_init with OMCGPU_E_UNSUPPORTED"""
    model, box, of = cases.build("mg_legendre")
    eng = lib.Engine(emu_lib)
    for field, bad in (("surf_type", 99), ("surf_bc", 7)):
        keep = int(model[field][0])
        model[field][0] = bad
        try:
            with pytest.raises(lib.OmcGpuError) as e:
                eng.init(model, 0, 100)
            assert _code(e) == abi.OMCGPU_E_UNSUPPORTED
        finally:
            model[field][0] = keep
    with pytest.raises(lib.OmcGpuError) as e:
        eng.init(model, 0, 0)
    assert _code(e) == abi.OMCGPU_E_INVALID_ARGUMENT
    eng.init(model, 0, 100)       # the untouched model is accepted
    eng.finalize()


def test_bounds_are_checked(emu_lib):
    model, box, of = cases.build("mg_temperature_multi")     # two cell-filter flux tallies
    src = Oracle(model).sample_initial_source(200, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 100)
    with pytest.raises(lib.OmcGpuError) as e:            # more sites than max_particles
        eng.set_source_bank(src)
    assert _code(e) == abi.OMCGPU_E_OUT_OF_BOUNDS
    eng.set_source_bank(src[:100])
    with pytest.raises(lib.OmcGpuError) as e:            # more histories than the bank holds
        eng.run_generation(eng.gen_in(150, 0, 0, 1.0, False))
    assert _code(e) == abi.OMCGPU_E_OUT_OF_BOUNDS
    with pytest.raises(lib.OmcGpuError) as e:            # no such tally
        eng.tally_results(17, np.zeros(4))
    assert _code(e) == abi.OMCGPU_E_OUT_OF_BOUNDS
    n = int(model["tally_result_ofs"][1] - model["tally_result_ofs"][0])
    with pytest.raises(lib.OmcGpuError) as e:            # wrong length of the VALUE array
        eng.tally_results(0, np.zeros(n + 1))
    assert _code(e) == abi.OMCGPU_E_INVALID_ARGUMENT
    idx = np.zeros(8, np.int32)
    cnt = C.c_int64(0)
    rc = eng.lib.omcgpu_comb_indices(C.c_int64(0), C.c_int64(1000), C.c_int64(1000), C.c_int64(900), C.c_int64(1),
                                     C.c_int64(1), C.c_uint64(152917), idx.ctypes.data_as(C.c_void_p),
                                     C.c_int64(idx.size), C.byref(cnt))
    assert rc == abi.OMCGPU_E_OUT_OF_BOUNDS and cnt.value == 900
    assert b"comb" in eng.lib.omcgpu_last_error()
    eng.finalize()


def test_host_bank_is_authoritative_between_generations(emu_lib):
    """what the host wrote into its source bank between two generations is what the next
    generation transports (omcgpu_transport_generation uploads the bank it is given)"""
    model, box, of = cases.build("mg_legendre")
    o = Oracle(model)
    src = o.sample_initial_source(400, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 400)
    gi = eng.gen_in(400, 0, 0, 1.0, False)
    _, out_a = eng.transport_generation(gi, src)
    edited = src.copy()
    edited["wgt"][:200] = 0.0                             # the host kills half of the sites
    _, out_b = eng.transport_generation(gi, edited)
    _, out_o = o.transport_generation(gi, edited)
    assert out_b.total_weight == 200.0 and out_a.total_weight == 400.0
    assert out_b.n_fission == out_o.n_fission and out_b.n_events == out_o.n_events
    assert out_b.n_events < out_a.n_events
    eng.finalize()


def test_integer_yield_secondaries_need_the_extended_build(emu_lib):
    """an inelastic product whose yield is an integer >= 2 makes the reference bank yield - 1
    secondary neutrons (src/physics.cpp:1158-1162).  The device transports them in the extended
    build (libomcgpu_ext.so; the host-compiled engine of this tier is that build):
    omcgpu_needs_extended_build names such a model, a non-integer yield only scales the weight"""
    import ctypes as C
    model, box, of, _ = cases.build_ce()
    rx = int(np.flatnonzero(np.asarray(model["rx_inelastic"]))[0])
    fn = int(model["prod_yield_fn"][int(model["rx_prod_ofs"][rx])])
    o0 = int(model["fn_ofs"][fn])
    assert int(model["fn_type"][fn]) == abi.OMCGPU_FN_POLYNOMIAL and int(model["fn_ofs"][fn + 1]) - o0 == 1
    keep = float(model["fn_data"][o0])
    eng = lib.Engine(emu_lib)
    try:
        assert eng.lib.omcgpu_needs_extended_build(model.byref()) == 0
        model["fn_data"][o0] = 2.0
        assert eng.lib.omcgpu_needs_extended_build(model.byref()) == 1
        eng.init(model, 0, 100)             # the extended build takes it
        eng.finalize()
        model["fn_data"][o0] = 1.5          # a non-integer yield only scales the weight
        assert eng.lib.omcgpu_needs_extended_build(model.byref()) == 0
        eng.init(model, 0, 100)
        eng.finalize()
    finally:
        model["fn_data"][o0] = keep


def test_full_fission_bank_stops_banking_like_the_reference(emu_lib, monkeypatch):
    """fission bank full: the reference warns, takes the progeny counter back and leaves the site
    loop (src/physics.cpp:233-247) -- no further random numbers are drawn for sites that cannot
    be stored.  The oracle does the same; with a bank of 3 * 40 sites and k_eff = 0.02 both must
    bank the same sites, and the count of histories that hit the full bank is reported.  (Which
    sites fit depends on the order of the collisions -- "results may be non-deterministic" says the
    reference -- so the histories run one after the other here, as in the oracle.)"""
    monkeypatch.setenv("OMCGPU_TAIL_THRESHOLD", "1000000")
    model, box, of = cases.build("mg_legendre")
    o = Oracle(model)
    n = 40
    src = o.sample_initial_source(n, box, of)
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 0.02, False)
    fb, out = eng.transport_generation(gi, src)
    fb_o, out_o = o.transport_generation(gi, src, 3 * n)
    assert out.n_fission == 3 * n == out_o.n_fission and out.n_bank_full > 0
    for f in ("E", "parent_id", "progeny_id", "delayed_group"):
        assert np.array_equal(fb[f], fb_o[f]), f
    assert out.n_events == out_o.n_events
    eng.finalize()


def test_update_reuploads_temperatures_and_densities(emu_lib):
    """omcgpu_update: what openmc_cell_set_temperature / openmc_material_set_densities change
    between batches (include/openmc/capi.h:30,98) reaches the device without a new omcgpu_init:
    a generation after the update equals the oracle's on the edited model"""
    model, box, of = cases.build("mg_temperature_multi")    # two cells at different temperatures
    o = Oracle(model)
    n = 600
    src = o.sample_initial_source(n, box, of)
    # (the two temperature blocks of the regression case's UO2 carry the same numbers: make the
    # hotter block absorb 5 % more, so that a cell's temperature shows in the transport)
    G = int(model["mg_n_groups"])
    ba_hot = int(model["blk_ang_ofs"][int(model["mgm_blk_ofs"][0]) + int(np.argmax(model["mgm_kTs"][:2]))])
    keep_abs = model["mg_absorption"].copy(); keep_tot = model["mg_total"].copy()
    model["mg_total"][ba_hot * G:(ba_hot + 1) * G] += 0.05 * keep_abs[ba_hot * G:(ba_hot + 1) * G]
    model["mg_absorption"][ba_hot * G:(ba_hot + 1) * G] *= 1.05
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    gi = eng.gen_in(n, 0, 0, 1.0, False)
    _, out0 = eng.transport_generation(gi, src)
    keep = model["cell_sqrtkT"].copy()
    try:
        model["cell_sqrtkT"][:] = keep.max()                # every cell at the hotter temperature
        eng.update(model, abi.OMCGPU_UPDATE_CELL_TEMPERATURES)
        _, out1 = eng.transport_generation(gi, src)
        _, out_o = Oracle(model).transport_generation(gi, src)
        assert out1.n_fission == out_o.n_fission and out1.n_events == out_o.n_events
        assert out1.n_events != out0.n_events              # the temperatures do matter in this model
        with pytest.raises(lib.OmcGpuError) as e:
            eng.update(model, 64)
        assert _code(e) == abi.OMCGPU_E_INVALID_ARGUMENT
    finally:
        model["cell_sqrtkT"][:] = keep
        model["mg_absorption"][:] = keep_abs
        model["mg_total"][:] = keep_tot
    eng.finalize()
    # densities, on a continuous-energy model
    model, box, of, _ = cases.build_ce()
    src = Oracle(model).sample_initial_source(n, box, of)
    eng.init(model, 0, n)
    keep = model["mat_atom_density"].copy()
    try:
        model["mat_atom_density"][:] = keep * 0.9
        eng.update(model, abi.OMCGPU_UPDATE_MATERIAL_DENSITIES)
        fb, out1 = eng.transport_generation(gi, src)
        fb_o, out_o = Oracle(model).transport_generation(gi, src)
        assert out1.n_fission == out_o.n_fission and out1.n_events == out_o.n_events
        assert np.array_equal(fb["E"], fb_o["E"])
    finally:
        model["mat_atom_density"][:] = keep
    eng.finalize()


def test_clean_source_bank_skips_the_upload(emu_lib):
    """gen_in.source_bank_clean = 1: the host vouches that it has not touched its bank since the
    library wrote it; the device-resident bank is used and the host buffer is not read"""
    model, box, of = cases.build("mg_legendre")
    o = Oracle(model)
    n = 400
    bank = np.ascontiguousarray(o.sample_initial_source(n, box, of))
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, n)
    out_a = eng.next_generation(eng.gen_in(n, 0, 0, 1.0, False), bank, 1)     # uploads, leaves the next source on both sides
    after = bank.copy()
    bank["wgt"][:] = 0.0                                                        # garbage the host copy ...
    out_b = eng.next_generation(eng.gen_in(n, 0, n, 1.0, False, source_bank_clean=True), bank, 2)
    ref = after.copy()
    out_c = o_next = None
    eng2 = lib.Engine(emu_lib)       # the same second generation from the intact host copy
    eng.finalize()
    eng2.init(model, 0, n)
    out_c = eng2.next_generation(eng2.gen_in(n, 0, n, 1.0, False), ref, 2)
    assert out_b.total_weight == n == out_c.total_weight                        # ... it was not read
    assert out_b.n_fission == out_c.n_fission and out_b.n_events == out_c.n_events
    assert np.array_equal(bank["E"], ref["E"])
    eng2.finalize()
