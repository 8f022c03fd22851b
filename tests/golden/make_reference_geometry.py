#!/usr/bin/env python
"""Writes tests/golden/refgeom_<config>.npz: probe points and what the REFERENCE's own
geometry code (oracle/_ref/libref_geom.so = src/surface.cpp, cell.cpp, lattice.cpp,
universe.cpp, geometry.cpp ... compiled unmodified from /root/reference) answers for
them.  Run here, where /root/reference exists; the GPU box only reads the vectors.

    python tests/golden/make_reference_geometry.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_geometry            # noqa: E402
import reference_cases as rc   # noqa: E402

for name in rc.NAMES:
    m, model, box = rc.build(name)
    ref = ref_geometry.ReferenceGeometry(m)
    xyz, uvw = rc.probe_points(box, 6000, seed=23)
    wx, wu = rc.walk(ref, xyz[:1000], uvw[:1000], steps=4)
    xyz, uvw = np.concatenate([xyz, wx]), np.concatenate([uvw, wu])
    cid, inst, mat, nco = ref.find_cell(xyz, uvw)
    d, s, lv, lt = ref.distance_to_boundary(xyz, uvw)
    out = os.path.join(HERE, f"refgeom_{name}.npz")
    np.savez_compressed(out, xyz=xyz, uvw=uvw, cell_id=cid, instance=inst, material=mat, n_coord=nco,
                        distance=d, surface=s, level=lv, lattice_trans=lt)
    print(name, xyz.shape[0], "points,", int(np.sum(cid >= 0)), "inside,", os.path.getsize(out) // 1024, "KiB")
