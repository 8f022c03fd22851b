"""The C-ABI shared library: it builds for sm_100a, loads, exports every symbol
include/omcgpu.h declares, and refuses to run without a CUDA device."""
import ctypes as C
import re

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
