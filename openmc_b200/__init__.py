"""openmc_b200 -- B200-native event-based transport engine behind OpenMC's C API.

Only what the hot path needs lives here:

* ``csrc/``        CUDA kernels for sm_100a + the C ABI of include/omcgpu.h
* ``lib``          ctypes binding of libomcgpu.so and the host-side eigenvalue driver
                   mirroring openmc.lib / src/simulation.cpp
* ``model``/``mgxs``/``nuclide``  host-side flattening of the read-only model
* ``examples``     the BASELINE.json configs as synthetic inputs
"""
__version__ = "0.1.0"
