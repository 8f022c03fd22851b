"""The five BASELINE.json geometries as (model description, flattened model, source box)
for the tests that put the reference's own geometry code next to the oracle / the CUDA
path, and the probe points they use."""
import numpy as np

import cases
from openmc_b200 import examples, mgxs, nuclide

NAMES = ["pincell_ce", "pincell_mg", "assembly_ce", "c5g7_core", "pwr_core_ce", "surface_zoo", "rotated_fills", "periodic_box"]
_CACHE = {}


def build(name):
    if name not in _CACHE:
        if name == "pincell_ce":          # C2
            m, lib, box, of = examples.pincell_ce(n_grid=400, tallies=False)
            model = nuclide.build_ce_model(m, lib)
        elif name == "pincell_mg":        # C1 geometry
            m, lib, box, of = examples.pincell_multigroup(mesh_dim=4, energy_bins=4)
            model = mgxs.build_mg_model(m, lib)
        elif name == "assembly_ce":       # C3
            m, lib, box, of = examples.assembly_ce(n_grid=400, tallies=False)
            model = nuclide.build_ce_model(m, lib)
        elif name == "c5g7_core":         # C4
            m, lib, box, of = examples.c5g7_core(mesh_dim=17, tallies=False)
            model = mgxs.build_mg_model(m, lib)
        elif name == "periodic_box":      # translational periodic planes (axis and oblique)
            m, lib, box, of = examples.periodic_box()
            model = mgxs.build_mg_model(m, lib)
        elif name == "rotated_fills":     # translated + rotated universe and lattice fills
            m, lib, box, of = examples.rotated_fills()
            model = mgxs.build_mg_model(m, lib)
        elif name == "surface_zoo":       # every supported surface kind, union / complement regions
            m, lib, box, of = examples.surface_zoo()
            model = mgxs.build_mg_model(m, lib)
        else:                             # C5
            m, lib, box, of = examples.pwr_core_ce(n_grid=300, n_fuel_nuclides=8, tallies=False)
            model = nuclide.build_ce_model(m, lib)
        _CACHE[name] = (m, model, box)
    return _CACHE[name]


def probe_points(box, n, seed=11):
    """random points / directions over (and a little beyond) the source box, points ON
    lattice / pin-cell planes (coincidence rules, include/openmc/geometry.h:33-36), and
    axis-parallel directions"""
    xyz, uvw = cases.geometry_probe_points(box, n, seed=seed)
    k = n // 10
    g = np.random.default_rng(seed + 1)
    axis = np.eye(3)[g.integers(0, 3, k)] * g.choice([-1.0, 1.0], size=(k, 1))
    uvw[2 * k:3 * k] = axis
    return xyz, uvw


def walk(ref, xyz, uvw, steps=6, nudge=1e-8):
    """points just past successive boundaries along each ray (the re-location problem of
    cross_surface, src/particle.cpp:549-629, posed to an exhaustive search)"""
    pts = []
    r = xyz.copy()
    for _ in range(steps):
        d = ref.distance_to_boundary(r, uvw)[0]
        ok = np.isfinite(d) & (d > 0) & (d < 1e6)
        r = np.where(ok[:, None], r + (d[:, None] + nudge) * uvw, r)
        pts.append(r.copy())
    return np.concatenate(pts), np.tile(uvw, (steps, 1))
