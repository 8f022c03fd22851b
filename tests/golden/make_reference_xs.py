#!/usr/bin/env python
"""Writes tests/golden/refxs_ce.npz: probe energies and what the REFERENCE's own
Nuclide::calculate_xs (oracle/_ref/libref_geom.so, src/nuclide.cpp compiled unmodified)
gives for the committed model tests/golden/ce_model.npz -- grid indices per nuclide and
density-weighted sums per material.  Run here, where /root/reference exists.

    python tests/golden/make_reference_xs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases            # noqa: E402
import ref_geometry     # noqa: E402

model = cases.load_model_npz(os.path.join(HERE, "ce_model.npz"))
ref = ref_geometry.ReferenceXS(model)
E = ref.in_range(cases.xs_probe_energies(model, None, 8000, seed=31))
out = {"E": E}
for mat in range(int(model["n_materials"])):
    if model["mat_nuc_ofs"][mat + 1] > model["mat_nuc_ofs"][mat]:
        out[f"xs{mat}"], out[f"ig{mat}"] = ref.macro(mat, E)
path = os.path.join(HERE, "refxs_ce.npz")
np.savez_compressed(path, **out)
print(E.size, "energies,", os.path.getsize(path) // 1024, "KiB")
