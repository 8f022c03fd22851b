#!/usr/bin/env python
"""Writes tests/golden/refcoll_ce.npz: neutrons about to collide and what the REFERENCE's own
collision() (oracle/_ref/libref_geom.so: src/physics.cpp, reaction.cpp, reaction_product.cpp,
physics_common.cpp compiled unmodified) leaves of them in the committed model
tests/golden/ce_model.npz -- energy, direction, weight, k-absorption score, seed afterwards,
sampled nuclide and MT, banked fission sites; analog absorption ("a_*") and survival biasing
from weight 0.3 ("s_*").  Run here, where /root/reference exists.

    python tests/golden/make_reference_collisions.py
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

N, SEED, K_EFF = 6000, 37, 1.1
model = cases.load_model_npz(os.path.join(HERE, "ce_model.npz"))
ref = ref_geometry.ReferenceXS(model)
mat, E, u, seeds = ref_geometry.collision_probes(ref, N, SEED)
out = {"material": mat, "E_in": E, "u_in": u, "seed_in": seeds, "k_eff": np.float64(K_EFF)}
for tag, survival, wgt0 in (("a", 0, 1.0), ("s", 1, 0.3)):
    ref.load_collision_physics(survival)
    r = ref.collisions(mat, E, u, seeds, K_EFF, wgt0)
    ref.unload_collision_physics()
    ref = ref_geometry.ReferenceXS(model)          # fresh nuclides for the next pass
    out[f"{tag}_wgt0"] = np.float64(wgt0)
    for k, v in r.items():
        out[f"{tag}_{k}"] = v
path = os.path.join(HERE, "refcoll_ce.npz")
np.savez_compressed(path, **out)
print(E.size, "collisions,", os.path.getsize(path) // 1024, "KiB")
