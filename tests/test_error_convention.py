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
