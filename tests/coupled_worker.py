"""A coupled-calculation driver as openmc.lib users write them, on the reference's own C API
(include/openmc/capi.h) of libopenmc_gpu.so: run two batches, change a material's density and a
cell's temperature, run the remaining batches; print k of every generation.

    [OPENMC_GPU=1 OMCGPU_LIB=...] python tests/coupled_worker.py <libopenmc_gpu.so>   (cwd = model directory)
"""
import ctypes as C
import os
import sys

if __name__ == "__main__":
    lib = C.CDLL(sys.argv[1])
    argv = (C.c_char_p * 1)(b"openmc")
    assert lib.openmc_init(1, argv, None) == 0
    assert lib.openmc_simulation_init() == 0
    status = C.c_int(0)
    ks = []
    lib.openmc_get_keff.argtypes = [C.POINTER(C.c_double * 2)]
    n = 0
    while True:
        assert lib.openmc_next_batch(C.byref(status)) == 0
        n += 1
        k = (C.c_double * 2)()
        lib.openmc_get_keff(C.byref(k))
        ks.append(k[0])
        if n == 2 and not os.environ.get("COUPLED_NO_CHANGE"):
            # fuel material (index 0): 5 % lighter; cell with index 0: 900 K
            lib.openmc_material_set_density.argtypes = [C.c_int32, C.c_double, C.c_char_p]
            rho = C.c_double(0.0)
            lib.openmc_material_get_density.argtypes = [C.c_int32, C.POINTER(C.c_double)]
            assert lib.openmc_material_get_density(0, C.byref(rho)) == 0
            assert lib.openmc_material_set_density(0, rho.value * 0.95, b"g/cm3") == 0
            lib.openmc_cell_set_temperature.argtypes = [C.c_int32, C.c_double, C.POINTER(C.c_int32), C.c_bool]
            assert lib.openmc_cell_set_temperature(0, 900.0, None, False) == 0
        if status.value != 0:
            break
    print("KEFF", " ".join(repr(x) for x in ks))
    lib.openmc_simulation_finalize()
    lib.openmc_finalize()
