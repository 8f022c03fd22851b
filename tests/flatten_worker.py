"""Run inside a directory holding model.xml + data: loads libopenmc_gpu.so (the reference's own
library with the host side of the B200 path compiled in), lets the REFERENCE read the model
(openmc_init), calls the C++ flattener and dumps every field of omcgpu_model to an .npz.

    python tests/flatten_worker.py <libopenmc_gpu.so> <out.npz>      (cwd = the model directory)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmc_b200 import _abi as abi  # noqa: E402

if __name__ == "__main__":
    lib = C.CDLL(sys.argv[1])
    argv = (C.c_char_p * 1)(b"openmc")
    rc = lib.openmc_init(1, argv, None)
    assert rc == 0, rc
    # the log-hash grids and data::energy_min / max are made by openmc_simulation_init
    # (initialize_data, src/simulation.cpp:682-761) -- where the hook calls the flattener
    rc = lib.openmc_simulation_init()
    assert rc == 0, rc
    lib.openmc_gpu_flatten_model.restype = C.POINTER(abi.Model)
    lib.openmc_gpu_model_array_length.restype = C.c_int64
    lib.openmc_gpu_model_array_length.argtypes = [C.c_char_p]
    p = lib.openmc_gpu_flatten_model()
    if not p:
        err = (C.c_char * 256).in_dll(lib, "openmc_err_msg").value.decode()
        np.savez(sys.argv[2], __error__=np.array(err))
        sys.exit(0)
    m = p.contents
    out = {}
    for name, (dt, is_ptr) in abi.model_field_types().items():
        if not is_ptr:
            out[name] = np.array(getattr(m, name))
            continue
        n = lib.openmc_gpu_model_array_length(name.encode())
        if n < 0:
            continue        # derived by omcgpu_init, not by the flattener
        ptr = getattr(m, name)
        out[name] = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy() if n else np.zeros(0, dt)
    np.savez(sys.argv[2], **out)
    lib.openmc_finalize()
