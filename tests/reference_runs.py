"""Eigenvalue cases run END TO END by the reference's own binary (oracle/_ref/openmc, the
unmodified sources against oracle/h5shim) and by the oracle / the CUDA path: shared by
tests/test_reference_end_to_end.py, tests/test_gpu_reference_runs.py and
tests/golden/make_reference_runs.py."""
import os
import shutil
import tempfile

import numpy as np

import cases
from openmc_b200 import examples, mgxs, nuclide
from openmc_b200.model import Filter, Tally

# name -> (particles, batches, inactive)
RUNS = {
    "pincell_ce": (1000, 10, 5),          # C2, small grids, mesh x energy tallies + collision estimator
    "pincell_ce_temperatures": (1000, 8, 3),   # library at 294 / 600 / 900 K, cells at 880 / 620 / 600 / 575 K
    "assembly_ce": (1000, 8, 3),          # C3, lattice, distribcell + collision-estimator tallies
    "pincell_mg": (1000, 8, 3),           # C1
    "c5g7_core": (1000, 8, 3),            # C4, lattice of lattices, mesh tally
    "pwr_core_ce": (600, 6, 2),           # C5 geometry, 40-nuclide fuel
    "surface_zoo": (1000, 6, 2),          # every surface kind, union / complement regions
    "rotated_fills": (1000, 6, 2),
    "periodic_box": (1000, 6, 2),
    "white_pincell": (1000, 6, 2),         # C1 pin cell with white (diffuse) boundaries
    "quadric_pin": (1000, 6, 2),           # general quadric surfaces
    "rotational_wedge": (1000, 6, 2),      # rotational periodic boundaries
    "pincell_ce_analog": (1000, 8, 3),     # analog estimator: material x energy and mesh tallies, all six scores
    "pincell_ce_analog_survival": (1000, 8, 3),   # the same with survival biasing
    "pincell_mg_analog": (1000, 8, 3),     # analog estimator, multigroup
    "pincell_mg_analog_survival": (1000, 8, 3),
    "hex_bundle_y": (1000, 6, 2),          # hexagonal lattice, orientation y, mesh + distribcell tallies
    "hex_bundle_x3d": (1000, 6, 2),        # orientation x, two axial levels
    "pincell_ce_density": (1000, 8, 3),    # cell densities (density multipliers) on the fuel rings and the moderator
    "assembly_ce_density": (1000, 6, 2),   # one density per INSTANCE of the fuel cell (distribcell offsets)
    "pincell_ce_n2n": (1000, 8, 3),        # (n,2n) reactions: secondary neutrons from integer yields
    "c5g7_core_density": (1000, 6, 2),     # multigroup: one density per INSTANCE of the UO2 pins' fuel cell
}
_CACHE = {}


def build(name):
    """-> (description, library, flattened model, box, only_fissionable, run_CE)"""
    if name in _CACHE:
        return _CACHE[name]
    if name == "pincell_ce":
        m, lib, box, of = examples.pincell_ce(n_grid=4000, mesh_dim=8, energy_bins=16, tallies=True)
        nf, nt = len(m.filters), len(m.tallies)
        m.filters.append(Filter(nf + 1, "material", [mat.id for mat in m.materials]))
        m.tallies.append(Tally(nt + 1, [nf + 1], ["flux", "total", "absorption", "fission", "nu-fission", "scatter"],
                               estimator="collision"))
        ce = True
    elif name in ("pincell_ce_analog", "pincell_ce_analog_survival", "pincell_mg_analog", "pincell_mg_analog_survival"):
        ce = "_ce_" in name
        if ce:
            m, lib, box, of = examples.pincell_ce(n_grid=3000, mesh_dim=6, energy_bins=10, tallies=True)
        else:
            m, lib, box, of = examples.pincell_multigroup(mesh_dim=6, energy_bins=8)
        six = ["flux", "total", "scatter", "absorption", "fission", "nu-fission"]
        nf, nt = len(m.filters), len(m.tallies)
        mesh_f = next(f for f in m.filters if f.type == "mesh")
        energy_f = next(f for f in m.filters if f.type == "energy")
        m.filters.append(Filter(nf + 1, "material", [mat.id for mat in m.materials]))
        m.tallies.append(Tally(nt + 1, [nf + 1, energy_f.id], six, estimator="analog"))
        m.tallies.append(Tally(nt + 2, [mesh_f.id], ["flux", "nu-fission", "scatter"], estimator="analog"))
        m.tallies.append(Tally(nt + 3, [nf + 1], six, estimator="collision"))
        # a scattering matrix: incoming x outgoing energy (EnergyoutFilter: the energy after the collision)
        eo = [1.0e-5, 0.625, 1.0e3, 1.0e5, 20.0e6] if ce else sorted(lib.group_edges)
        m.filters.append(Filter(nf + 2, "energy", eo))
        m.filters.append(Filter(nf + 3, "energyout", eo))
        m.tallies.append(Tally(nt + 4, [nf + 2, nf + 3], ["scatter", "nu-fission"], estimator="analog"))
        if name.endswith("_survival"):
            m.settings["survival_biasing"] = 1
    elif name == "pincell_ce_density":
        m, lib, box, of = examples.pincell_ce(n_grid=3000, mesh_dim=6, energy_bins=12, tallies=True,
                                               temperatures=((294.0, 600.0, 900.0), (880.0, 620.0, 600.0, 575.0)))
        m.settings["temperature_tolerance"] = 50.0
        mats = [c for c in m.cells if c.fill_type == "material" and c.fill is not None]
        mats[0].density = 9.3          # inner fuel ring: below the material's density
        mats[1].density = 10.9         # outer ring: above
        mats[-1].density = 0.66        # moderator
        ce = True
    elif name == "assembly_ce_density":
        m, lib, box, of = examples.assembly_ce(n_grid=2000)
        m.settings["material_cell_offsets"] = 1
        fuel_cell = next(c for c in m.cells if c.fill_type == "material" and c.fill is not None
                         and any(n == "U235" for n, _ in c.fill.components))
        n_inst = 17 * 17 - 25
        fuel_cell.density = list(10.0 + 0.4 * np.cos(0.37 * np.arange(n_inst)))
        ce = True
    elif name == "pincell_ce_n2n":
        library = nuclide.synthetic_library(nuclide.PINCELL_NUCLIDES, 3000, n2n=True)
        m, lib, box, of = examples.pincell_ce(n_grid=3000, mesh_dim=6, energy_bins=12, tallies=True, library=library)
        ce = True
    elif name == "c5g7_core_density":
        m, lib, box, of = examples.c5g7_core(mesh_dim=17)
        m.settings["material_cell_offsets"] = 1
        model0 = mgxs.build_mg_model(m, lib)          # (for the number of instances of the cell)
        fuel = m.cells[0]
        n_inst = int(m._n_instances[0])
        fuel.density = list(1.6749 * (1.0 + 0.15 * np.cos(0.11 * np.arange(n_inst))))
        m.cells[1].density = 1.5                       # the water around those pins: one value
        ce = False
    elif name == "hex_bundle_y":
        m, lib, box, of = examples.hex_bundle("y")
        ce = False
    elif name == "hex_bundle_x3d":
        m, lib, box, of = examples.hex_bundle("x", n_axial=2)
        ce = False
    elif name == "pincell_ce_temperatures":
        m, lib, box, of = examples.pincell_ce(n_grid=3000, mesh_dim=6, energy_bins=12, tallies=True,
                                               temperatures=((294.0, 600.0, 900.0), (880.0, 620.0, 600.0, 575.0)))
        m.settings["temperature_tolerance"] = 50.0
        ce = True
    elif name == "assembly_ce":
        m, lib, box, of = examples.assembly_ce(n_grid=2000)
        nf, nt = len(m.filters), len(m.tallies)
        fuel_cell = next(c for c in m.cells if c.fill_type == "material" and c.fill is not None
                         and any(n == "U235" for n, _ in c.fill.components))
        m.filters.append(Filter(nf + 1, "distribcell", [fuel_cell.id]))
        m.tallies.append(Tally(nt + 1, [nf + 1], ["flux", "fission"]))
        m.tallies.append(Tally(nt + 2, [nf + 1], ["flux", "absorption"], estimator="collision"))
        ce = True
    elif name == "pwr_core_ce":
        m, lib, box, of = examples.pwr_core_ce(n_grid=600, n_fuel_nuclides=40, tallies=True)
        ce = True
    elif name == "pincell_mg":
        m, lib, box, of = examples.pincell_multigroup(mesh_dim=10, energy_bins=20)
        ce = False
    elif name == "c5g7_core":
        m, lib, box, of = examples.c5g7_core(mesh_dim=51)
        ce = False
    elif name == "white_pincell":
        m, lib, box, of = examples.pincell_multigroup(mesh_dim=6, energy_bins=8)
        for s in m.surfaces:
            if s.boundary == "reflective":
                s.boundary = "white"
        ce = False
    else:
        m, lib, box, of = getattr(examples, name)()
        ce = False
    model = nuclide.build_ce_model(m, lib) if ce else mgxs.build_mg_model(m, lib)
    _CACHE[name] = (m, lib, model, box, of, ce)
    return _CACHE[name]


def write_inputs(directory, m, lib, box, of, ce, particles, batches, inactive, **kw):
    from oracle import refrun
    if ce:
        refrun.write_ce_library(os.path.join(directory, "data"), lib)
        xs = "data/cross_sections.xml"
    else:
        refrun.write_mgxs_library(os.path.join(directory, "mgxs.h5"), lib)
        xs = "mgxs.h5"
    with open(os.path.join(directory, "model.xml"), "w") as f:
        f.write(refrun.model_xml(m, xs, particles, batches, inactive, box, of, run_CE=ce, **kw))


def digest(r):
    """what the tests compare, from a refrun.run() result: k per generation, k-combined,
    global tallies [4][sum, sum_sq], the tallies' [bins][scores][sum, sum_sq] concatenated in
    tally order, the final source bank"""
    sp = r["statepoint"]
    names = sorted((k for k in sp["tallies"].keys() if k.startswith("tally ")), key=lambda s: int(s.split()[1]))
    res = [np.asarray(sp["tallies"][k]["results"][()], float).reshape(-1, 2) for k in names]
    sb = sp["source_bank"][()]
    return dict(k_generation=np.array(sp["k_generation"][()], float), k_combined=np.array(sp["k_combined"][()], float),
                global_tallies=np.array(sp["global_tallies"][()], float)[:, 1:],
                tally_sum=np.concatenate([x[:, 0] for x in res]) if res else np.zeros(0),
                tally_sum_sq=np.concatenate([x[:, 1] for x in res]) if res else np.zeros(0),
                bank_r=np.stack([sb["r"]["x"], sb["r"]["y"], sb["r"]["z"]], axis=1),
                bank_u=np.stack([sb["u"]["x"], sb["u"]["y"], sb["u"]["z"]], axis=1),
                bank_E=np.array(sb["E"]), bank_wgt=np.array(sb["wgt"]),
                bank_delayed_group=np.array(sb["delayed_group"]))


def run_reference(name, threads=1, event=False):
    from oracle import refrun
    m, lib, model, box, of, ce = build(name)
    particles, batches, inactive = RUNS[name]
    d = tempfile.mkdtemp(prefix="omc_e2e_")
    try:
        write_inputs(d, m, lib, box, of, ce, particles, batches, inactive)
        r = refrun.run(d, threads=threads, event=event)
        assert r["returncode"] == 0, r["stdout"][-1500:] + r["stderr"][-1500:]
        return digest(r)
    finally:
        shutil.rmtree(d, ignore_errors=True)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refrun_%s.npz")
