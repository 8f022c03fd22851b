"""Shared model builders for the tests: the reference's self-contained MG
regression cases with their golden results (results_true.dat)."""
from openmc_b200 import examples, mgxs

# name -> (builder, golden k-combined mean, std, golden tallies [(sum, sum_sq), ...])
GOLDEN = {
    # tests/regression_tests/mg_legendre/results_true.dat
    "mg_legendre": ("1.009220E+00", "9.571832E-03", []),
    # tests/regression_tests/mg_max_order/results_true.dat
    "mg_max_order": ("1.009220E+00", "9.571832E-03", []),
    # tests/regression_tests/mg_survival_biasing/results_true.dat
    "mg_survival_biasing": ("9.889968E-01", "9.144186E-03", []),
    # tests/regression_tests/mg_basic/results_true.dat
    "mg_basic": ("1.004679E+00", "1.329350E-02", []),
    # tests/regression_tests/mg_temperature_multi/results_true.dat
    "mg_temperature_multi": ("1.309371E+00", "6.765039E-03",
                             [("2.532303E+01", "1.282689E+02"), ("9.336894E+01", "1.743765E+03")]),
}


def build(name):
    """-> (ModelHandle, source box, only_fissionable); 1000 particles, 10 batches, 5 inactive."""
    if name == "mg_legendre":
        m, box, of = examples.slab_mg()
        return mgxs.build_mg_model(m, examples.slab_2g_library(1)), box, of
    if name == "mg_max_order":
        m, box, of = examples.slab_mg()
        m.settings["max_order"] = 1
        return mgxs.build_mg_model(m, examples.slab_2g_library(2)), box, of
    if name == "mg_survival_biasing":
        m, box, of = examples.slab_mg()
        m.settings["survival_biasing"] = 1
        return mgxs.build_mg_model(m, examples.slab_2g_library(1)), box, of
    if name == "mg_basic":
        m, lib, box, of = examples.mg_basic()
        return mgxs.build_mg_model(m, lib), box, of
    if name == "mg_temperature_multi":
        m, lib, box, of = examples.mg_temperature_multi()
        return mgxs.build_mg_model(m, lib), box, of
    raise KeyError(name)


def fmt(x):
    return "%.6E" % x


_CE_CACHE = {}


def build_ce(n_grid=4000, tallies=True, mesh_dim=8, energy_bins=16):
    """small synthetic CE pincell (config C2 geometry, 10 nuclides) for parity tests"""
    from openmc_b200 import nuclide
    key = (n_grid, tallies, mesh_dim, energy_bins)
    if key not in _CE_CACHE:
        m, libx, box, of = examples.pincell_ce(n_grid=n_grid, mesh_dim=mesh_dim,
                                               energy_bins=energy_bins, tallies=tallies)
        _CE_CACHE[key] = (nuclide.build_ce_model(m, libx), box, of, libx)
    return _CE_CACHE[key]


def xs_probe_energies(model, libx, n, seed=7):
    """log-uniform energies + the edge cases of SURVEY Appendix B.1: exact grid
    points, first/last points, out-of-range values, hash-bin edges."""
    import numpy as np
    g = np.random.default_rng(seed)
    e_min, e_max = float(model["energy_min"]), float(model["energy_max"])
    E = np.exp(g.uniform(np.log(e_min), np.log(e_max), n))
    grid = np.asarray(model["nuc_energy"])
    pick = grid[g.integers(0, grid.size, n // 10)]
    spacing = float(model["log_spacing"])
    edges = e_min * np.exp(spacing * g.integers(1, int(model["n_log_bins"]), n // 10))
    near = np.concatenate([np.nextafter(edges, 0), edges, np.nextafter(edges, np.inf)])
    # edges of the device's fine hash (OMCGPU_FINE_PER_LOG_BIN bins per reference bin)
    from openmc_b200 import _abi as abi
    per = abi.OMCGPU_FINE_PER_LOG_BIN
    fine = e_min * np.exp(spacing / per * g.integers(1, int(model["n_log_bins"]) * per, n // 10))
    near = np.concatenate([near, np.nextafter(fine, 0), fine, np.nextafter(fine, np.inf)])
    special = np.array([e_min, e_max, e_min * 0.5, e_max * 1.5, np.nextafter(e_min, 1), np.nextafter(e_max, 0)])
    return np.concatenate([E, pick, np.nextafter(pick, 0), np.nextafter(pick, np.inf), near, special])


def save_model_npz(path, model):
    """every omcgpu_model field (arrays and scalars) of a ModelHandle -> one .npz"""
    import numpy as np
    from openmc_b200 import _abi as abi
    out = {}
    for name, (dt, is_ptr) in abi.model_field_types().items():
        out[name] = model[name] if is_ptr else np.array(model[name])
    np.savez_compressed(path, **out)


def load_model_npz(path):
    import numpy as np
    from openmc_b200 import _abi as abi
    z = np.load(path)
    types = abi.model_field_types()
    vals = {k: (z[k] if types[k][1] else z[k].item()) for k in z.files if k in types}
    # per-cell fields added to the ABI after a fixture was written: no rotated fills
    nc = int(vals["n_cells"])
    vals.setdefault("cell_rotated", np.zeros(nc, np.int32))
    vals.setdefault("cell_rotation", np.zeros(9 * nc))
    # per-tally estimator added later: the fixtures' tallies are tracklength tallies
    vals.setdefault("tally_estimator", np.full(max(int(vals["n_tallies"]), 1), abi.OMCGPU_ESTIMATOR_TRACKLENGTH, np.int32))
    return abi.ModelHandle(vals)


def build_lattice(name):
    """lattice configurations: 'assembly' = config C3 (17 x 17, CE, 2 levels),
    'c5g7' = config C4 (lattice of lattices, MG, 3 levels), 'pwr_core' = config C5
    (Hoogenboom-Martin full core, CE, 3 levels, z-planes); small grids / meshes"""
    from openmc_b200 import nuclide
    key = ("lattice", name)
    if key not in _CE_CACHE:
        if name == "assembly":
            m, libx, box, of = examples.assembly_ce(n_grid=2000)
            _CE_CACHE[key] = (nuclide.build_ce_model(m, libx), box, of)
        elif name == "pwr_core":     # config C5 geometry, 40-nuclide fuel, small grids
            m, libx, box, of = examples.pwr_core_ce(n_grid=600, n_fuel_nuclides=40, tallies=True)
            _CE_CACHE[key] = (nuclide.build_ce_model(m, libx), box, of)
        else:
            m, lib, box, of = examples.c5g7_core(mesh_dim=51)
            _CE_CACHE[key] = (mgxs.build_mg_model(m, lib), box, of)
    return _CE_CACHE[key]


def geometry_probe_points(box, n, seed=4, margin=0.05):
    """random points / directions over the source box (a little beyond it), plus
    points placed exactly on lattice / pin-cell boundaries (coincidence rules)"""
    import numpy as np
    g = np.random.default_rng(seed)
    box = np.asarray(box, float)
    lo, hi = box[:3] - margin * (box[3:] - box[:3]), box[3:] + margin * (box[3:] - box[:3])
    xyz = lo + g.uniform(size=(n, 3)) * (hi - lo)
    uvw = g.normal(size=(n, 3))
    uvw /= np.linalg.norm(uvw, axis=1)[:, None]
    k = n // 10
    pitch = 1.26
    xyz[:k, 0] = np.round(xyz[:k, 0] / pitch) * pitch          # on x lattice planes
    xyz[k:2 * k, 1] = np.round(xyz[k:2 * k, 1] / pitch) * pitch  # on y lattice planes
    return xyz, uvw


def bank_parity(fb, fb_o, lo, hi, min_same=0.998):
    """sorted fission banks of the histories (lo, hi] compared parent by parent: `fb` may hold
    more parents (a whole generation), `fb_o` holds exactly those.  -> (fraction of parents
    with the same number of sites, fraction of their sites that are equal)"""
    import numpy as np
    sel = (fb["parent_id"] > lo) & (fb["parent_id"] <= hi)
    g = fb[sel]
    # per parent: number of sites
    cg = np.bincount(g["parent_id"] - lo - 1, minlength=hi - lo)
    co = np.bincount(fb_o["parent_id"] - lo - 1, minlength=hi - lo)
    same_count = cg == co
    # a last-ulp difference between the device libm and glibc may flip a rare discrete
    # decision of a history (oracle header); the other histories must agree site by site
    assert same_count.mean() >= min_same, same_count.mean()
    keep_g = same_count[g["parent_id"] - lo - 1]
    keep_o = same_count[fb_o["parent_id"] - lo - 1]
    a, b = g[keep_g], fb_o[keep_o]
    assert a.size == b.size and a.size > 0.9 * fb_o.size
    for f in ("parent_id", "progeny_id"):
        assert np.array_equal(a[f], b[f]), f
    dg_same = a["delayed_group"] == b["delayed_group"]
    close = np.isclose(a["E"], b["E"], rtol=1e-9, atol=0) & dg_same & \
        np.all(np.isclose(a["r"], b["r"], rtol=0, atol=1e-7), axis=1)
    assert close.mean() >= min_same, close.mean()
    return float(same_count.mean()), float(close.mean())
