"""Mesh ray trace of the tracklength tallies against the REFERENCE ITSELF:
RegularMesh::bins_crossed (StructuredMesh::raytrace_mesh, src/mesh.cpp:999-1146, compiled
unmodified into oracle/_ref/libref_geom.so) for tracks inside, through, grazing and outside
1-, 2- and 3-dimensional meshes: the oracle must emit the same bins with bit-identical length
fractions."""
import os

import numpy as np
import pytest

import ref_geometry
from openmc_b200 import examples, mgxs, nuclide
from oracle.binding import Oracle

needs_ref = pytest.mark.skipif(not os.path.exists(ref_geometry.LIB),
                               reason="oracle/_ref/libref_geom.so not built (no /root/reference here)")


def _tracks(mesh, n, seed):
    g = np.random.default_rng(seed)
    nd = len(mesh.dimension)
    ll = np.zeros(3); ur = np.ones(3)
    ll[:nd], ur[:nd] = mesh.lower_left, mesh.upper_right
    if nd < 3:
        ll[nd:], ur[nd:] = -1.0, 1.0
    span = ur - ll
    r0 = ll - 0.3 * span + g.uniform(size=(n, 3)) * 1.6 * span       # inside and outside the mesh
    u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    k = n // 8
    u[:k] = np.eye(3)[g.integers(0, 3, k)] * g.choice([-1.0, 1.0], size=(k, 1))      # axis-parallel
    w = span / np.array(list(mesh.dimension) + [1] * (3 - nd), float)
    r0[k:2 * k, 0] = ll[0] + np.round((r0[k:2 * k, 0] - ll[0]) / w[0]) * w[0]        # starts ON mesh planes
    d = np.concatenate([g.exponential(0.3, n // 2), g.exponential(3.0, n - n // 2)]) * span.min()
    d[2 * k:2 * k + 50] = 1e-9                                                         # tiny tracks
    return r0, r0 + d[:, None] * u, u


@needs_ref
@pytest.mark.parametrize("case", ["pincell_ce", "assembly_ce", "c5g7_core", "slab_1d", "box_3d"])
def test_oracle_mesh_ray_trace_equals_reference(case):
    if case == "pincell_ce":
        m, lib, box, of = examples.pincell_ce(n_grid=300, mesh_dim=100)
        model = nuclide.build_ce_model(m, lib)
    elif case == "assembly_ce":
        m, lib, box, of = examples.assembly_ce(n_grid=300)
        model = nuclide.build_ce_model(m, lib)
    elif case == "c5g7_core":
        m, lib, box, of = examples.c5g7_core(mesh_dim=51)
        model = mgxs.build_mg_model(m, lib)
    else:
        m, lib, box, of = examples.pincell_multigroup(mesh_dim=4, energy_bins=4)
        from openmc_b200.model import Mesh
        if case == "slab_1d":
            m.meshes[0] = Mesh(id=m.meshes[0].id, dimension=[37], lower_left=[-0.63], upper_right=[0.63])
        else:
            m.meshes[0] = Mesh(id=m.meshes[0].id, dimension=[7, 5, 3], lower_left=[-0.63, -0.63, -1.0],
                               upper_right=[0.63, 0.63, 1.0])
        model = mgxs.build_mg_model(m, lib)
    ref = ref_geometry.ReferenceMesh(m.meshes)
    o = Oracle(model)
    for i, mesh in enumerate(m.meshes):
        r0, r1, u = _tracks(mesh, 4000, seed=5 + i)
        ofs_r, bins_r, len_r = ref.bins_crossed(i, r0, r1, u)
        ofs_o, bins_o, len_o = o.mesh_bins_crossed(i, r0, r1, u)
        assert bins_r.size > 4000
        assert np.array_equal(ofs_o, ofs_r) and np.array_equal(bins_o, bins_r)
        assert np.array_equal(len_o, len_r)
