"""The C-ABI shared library: it builds for sm_100a, loads, exports every symbol
include/omcgpu.h declares, and refuses to run without a CUDA device."""
import ctypes as C
import re

import numpy as np

import pytest

from openmc_b200 import _abi as abi
from openmc_b200 import lib
from openmc_b200.csrc import build


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(abi.HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(omcgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_structs():
    assert C.sizeof(abi.Site) == 104      # SourceSite, include/openmc/particle_data.h:43-57
    assert abi.site_dtype.itemsize == 104
    assert C.sizeof(abi.GenIn) == 56
    names = [f for f, *_ in abi.STRUCTS["omcgpu_model"]]
    assert len(names) == len(set(names)) and "nuc_xs" in names and "sc_dist" in names


def test_library_builds_and_exports_every_symbol():
    path = build.build()
    L = C.CDLL(path)
    fns = declared_functions()
    assert len(fns) >= 20
    for fn in fns:
        assert hasattr(L, fn), fn


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cases
    model, _, _ = cases.build("mg_legendre")
    eng = lib.Engine()
    with pytest.raises(lib.OmcGpuError, match="no CUDA device"):
        eng.init(model, 0, 1000)


def test_which_models_need_the_extended_build():
    """omcgpu_needs_extended_build (pure host code of the PRODUCT library, no device needed): the
    models with hexagonal lattices, axis cones / general quadrics, white / rotational periodic
    boundaries, analog-estimator tallies or secondary neutrons of integer yields are the ones the
    loaders hand to libomcgpu_ext.so; everything else -- BASELINE.json's five configurations
    among them -- runs on the default build, whose kernels carry none of that code"""
    import ctypes as C
    import reference_runs as rr
    from openmc_b200 import lib as omclib
    L = C.CDLL(omclib.DEFAULT_LIB)
    extended = {"surface_zoo", "white_pincell", "quadric_pin", "rotational_wedge", "pincell_ce_analog",
                "pincell_ce_analog_survival", "pincell_mg_analog", "pincell_mg_analog_survival", "hex_bundle_y",
                "hex_bundle_x3d", "pincell_ce_n2n", "c5g7_core_density"}
    for name in rr.RUNS:
        if name in ("pwr_core_ce", "assembly_ce_density"):      # (slow to build; nothing new in them)
            continue
        m, libx, model, box, of, ce = rr.build(name)
        assert bool(L.omcgpu_needs_extended_build(model.byref())) == (name in extended), name


def test_descriptions_flattened_before_a_field_existed_keep_their_meaning():
    """tests/golden/*.npz were flattened before lat_type, cell_density_*, surf_periodic_axis and
    tally_estimator joined omcgpu_model: ModelHandle fills in what their absence meant
    (rectangular lattices, no density multipliers, translational periodicity, tracklength
    tallies) with arrays of the model's own sizes -- not one zero the library would read past"""
    m, libx, model, box, of, ce = __import__("reference_runs").build("c5g7_core")
    values = {k: model[k] for k, *_ in abi.STRUCTS["omcgpu_model"]}
    for k in ("lat_type", "cell_density_ofs", "cell_density_mult", "surf_periodic_axis", "tally_estimator"):
        values.pop(k)
    old = abi.ModelHandle(values)
    nl, nc, ns, nt = (int(model[k]) for k in ("n_lattices", "n_cells", "n_surfaces", "n_tallies"))
    assert old["lat_type"].size == nl and not old["lat_type"].any()
    assert np.array_equal(old["cell_density_ofs"], np.arange(nc + 1)) and np.all(old["cell_density_mult"] == 1.0)
    assert old["surf_periodic_axis"].size == ns and np.all(old["surf_periodic_axis"] == -1)
    assert old["tally_estimator"].size == nt and np.all(old["tally_estimator"] == abi.OMCGPU_ESTIMATOR_TRACKLENGTH)
    import ctypes as C
    from openmc_b200 import lib as omclib
    assert C.CDLL(omclib.DEFAULT_LIB).omcgpu_needs_extended_build(old.byref()) == 0
