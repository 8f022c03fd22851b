"""Run the UNMODIFIED reference (oracle/_ref/openmc, built by oracle/Makefile from
/root/reference against oracle/h5shim) on a model of this repo -- TEST / BASELINE
INFRASTRUCTURE: the end-to-end oracle of the CE / MG configurations and the CPU arm of
bench.py (`cpu_baseline.kind = "reference"`).

* ``model_xml``           the reference's single-file input (model.xml: materials, geometry,
                          settings, tallies; formats docs/source/io_formats/*.rst, read by
                          src/initialize.cpp read_model_xml)
* ``write_ce_library``    one container file per synthetic nuclide in the layout of
                          docs/source/io_formats/nuclear_data.rst (writer the reference ships:
                          openmc/data/neutron.py:354-431) + cross_sections.xml
                          (src/cross_sections.cpp:45-88)
* ``write_mgxs_library``  mgxs.h5 in the layout of docs/source/io_formats/mgxs_library.rst
                          (writer: openmc/mgxs_library.py:1952-2146, 2504-2529)
* ``run``                 executes the binary (OpenMP threads, history-based or --event) and
                          returns k-effective, the rates its own output prints
                          (src/output.cpp:424-497) and the statepoint tree

The files are OMH5 containers (oracle/omh5.py), which is what the reference built with
oracle/h5shim reads through its HDF5 calls; the tree is the documented one.
"""
from __future__ import annotations

import os
import re
import subprocess

import numpy as np

from . import omh5

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(HERE, "_ref", "openmc")
K_BOLTZMANN = 8.617333262e-5   # include/openmc/constants.h

LATTICE_ID_BASE = 100000   # universes and lattices share one id space in geometry.xml
                           # (a cell's fill is looked up as a universe first,
                           # src/geometry_aux.cpp:90-106); this repo's description refers to
                           # lattices by object


def available():
    return os.path.exists(BINARY)


def _fmt(x):
    return " ".join(repr(float(v)) for v in x)


# ---------------------------------------------------------------------------------------------
# XML
# ---------------------------------------------------------------------------------------------
def geometry_elements(m) -> list:
    out = []
    for s in m.surfaces:
        bc = "" if s.boundary in ("transmission", None) else f' boundary="{s.boundary}"'
        if getattr(s, "periodic_surface_id", None) is not None:
            bc += f' periodic_surface_id="{s.periodic_surface_id}"'
        out.append(f'  <surface id="{s.id}" type="{s.kind}" coeffs="{_fmt(s.coeffs)}"{bc}/>')
    for c in m.cells:
        a = [f'id="{c.id}"', f'universe="{c.universe}"']
        if c.fill_type == "material":
            a.append('material="void"' if c.fill is None else f'material="{c.fill.id}"')
        elif c.fill_type == "universe":
            a.append(f'fill="{c.fill}"')
        else:
            a.append(f'fill="{LATTICE_ID_BASE + c.fill.id}"')
        if c.region:
            a.append(f'region="{c.region}"')
        if getattr(c, "temperature", None) is not None and c.fill_type == "material":
            a.append(f'temperature="{float(c.temperature)!r}"')
        if getattr(c, "density", None) is not None:
            a.append(f'density="{_fmt(np.atleast_1d(np.asarray(c.density, float)))}"')
        if any(float(t) != 0.0 for t in c.translation):
            a.append(f'translation="{_fmt(c.translation)}"')
        if getattr(c, "rotation", None) is not None:
            a.append(f'rotation="{_fmt(c.rotation)}"')
        out.append("  <cell " + " ".join(a) + "/>")
    for l in m.lattices:
        if hasattr(l, "n_rings"):      # HexLattice: <universes> in the order fill_lattice_x / _y walk it
            n = 2 * l.n_rings - 1
            arr = np.asarray(l.universes, dtype=np.int64).reshape(-1, n, n)
            out.append(f'  <hex_lattice id="{LATTICE_ID_BASE + l.id}" n_rings="{l.n_rings}"' +
                       (f' n_axial="{l.n_axial}"' if l.n_axial is not None else "") +
                       f' orientation="{l.orientation}">')
            out.append(f"    <pitch>{_fmt(l.pitch)}</pitch>")
            out.append(f"    <center>{_fmt(l.center)}</center>")
            if l.outer is not None:
                out.append(f"    <outer>{l.outer}</outer>")
            order = l.input_order()
            rows = [" ".join(str(int(arr[z, i1, i0])) for i1, i0 in order) for z in range(arr.shape[0])]
            out.append("    <universes>\n" + "\n".join(rows) + "\n    </universes>")
            out.append("  </hex_lattice>")
            continue
        arr = np.asarray(l.universes)
        if arr.ndim == 2:
            ny, nx = arr.shape
            dim, nd, slabs = f"{nx} {ny}", 2, [arr]
        else:
            nz, ny, nx = arr.shape
            dim, nd, slabs = f"{nx} {ny} {nz}", 3, list(arr)
        out.append(f'  <lattice id="{LATTICE_ID_BASE + l.id}">')
        out.append(f"    <pitch>{_fmt(l.pitch[:nd])}</pitch>")
        out.append(f"    <dimension>{dim}</dimension>")
        out.append(f"    <lower_left>{_fmt(l.lower_left[:nd])}</lower_left>")
        if l.outer is not None:
            out.append(f"    <outer>{l.outer}</outer>")
        rows = []
        for slab in slabs:
            rows += [" ".join(str(int(u)) for u in row) for row in slab]
        out.append("    <universes>\n" + "\n".join(rows) + "\n    </universes>")
        out.append("  </lattice>")
    return out


def geometry_xml(m) -> str:
    return "\n".join(["<?xml version='1.0' encoding='utf-8'?>", "<geometry>"] + geometry_elements(m) +
                     ["</geometry>"])


def materials_elements(m, cross_sections) -> list:
    out = [f"  <cross_sections>{cross_sections}</cross_sections>"]
    for mat in m.materials:
        t = "" if mat.temperature is None else f' temperature="{float(mat.temperature)!r}"'
        out.append(f'  <material id="{mat.id}"{t}>')
        if mat.units == "sum":
            out.append('    <density units="sum"/>')
        elif mat.units == "macro":
            out.append(f'    <density value="{1.0 if mat.density is None else float(mat.density)!r}" units="macro"/>')
        else:
            out.append(f'    <density value="{float(mat.density)!r}" units="{mat.units}"/>')
        if mat.macroscopic is not None:
            out.append(f'    <macroscopic name="{mat.macroscopic}"/>')
        else:
            for name, frac in mat.components:
                key = "ao" if frac >= 0 else "wo"
                out.append(f'    <nuclide name="{name}" {key}="{abs(float(frac))!r}"/>')
        out.append("  </material>")
    return out


def settings_elements(m, particles, batches, inactive, box, only_fissionable, run_CE, gen_per_batch=1,
                      seed=None, extra=None) -> list:
    st = m.settings
    out = ["  <run_mode>eigenvalue</run_mode>", f"  <particles>{int(particles)}</particles>",
           f"  <batches>{int(batches)}</batches>", f"  <inactive>{int(inactive)}</inactive>"]
    if gen_per_batch != 1:
        out.append(f"  <generations_per_batch>{int(gen_per_batch)}</generations_per_batch>")
    out += ['  <source type="independent" strength="1.0" particle="neutron">', '    <space type="box">',
            f"      <parameters>{_fmt(box)}</parameters>", "    </space>"]
    if only_fissionable:
        out += ["    <constraints>", "      <fissionable>true</fissionable>", "    </constraints>"]
    out.append("  </source>")
    if not run_CE:
        out.append("  <energy_mode>multi-group</energy_mode>")
    if seed is not None:
        out.append(f"  <seed>{int(seed)}</seed>")
    if st.get("survival_biasing"):
        out.append("  <survival_biasing>true</survival_biasing>")
    cut = []
    if st.get("weight_cutoff", 0.25) != 0.25:
        cut.append(f"    <weight>{st['weight_cutoff']!r}</weight>")
    if st.get("weight_survive", 1.0) != 1.0:
        cut.append(f"    <weight_avg>{st['weight_survive']!r}</weight_avg>")
    if st.get("energy_cutoff", 0.0) != 0.0:
        cut.append(f"    <energy_neutron>{st['energy_cutoff']!r}</energy_neutron>")
    if cut:
        out += ["  <cutoff>"] + cut + ["  </cutoff>"]
    if st.get("max_order") is not None:
        out.append(f"  <max_order>{int(st['max_order'])}</max_order>")
    if not st.get("legendre_to_tabular", True):
        out.append("  <tabular_legendre><enable>false</enable></tabular_legendre>")
    elif st.get("legendre_to_tabular_points") is not None:
        out.append("  <tabular_legendre><enable>true</enable><num_points>%d</num_points></tabular_legendre>"
                   % int(st["legendre_to_tabular_points"]))
    if st.get("n_log_bins", 8000) != 8000:
        out.append(f"  <log_grid_bins>{int(st['n_log_bins'])}</log_grid_bins>")
    if not st.get("create_delayed_neutrons", 1):
        out.append("  <create_delayed_neutrons>false</create_delayed_neutrons>")
    if st.get("temperature_tolerance", 10.0) != 10.0:
        out.append(f"  <temperature_tolerance>{st['temperature_tolerance']!r}</temperature_tolerance>")
    if st.get("temperature_method"):
        out.append(f"  <temperature_method>{st['temperature_method']}</temperature_method>")
    if st.get("temperature_default") is not None:
        out.append(f"  <temperature_default>{float(st['temperature_default'])!r}</temperature_default>")
    if st.get("material_cell_offsets") is False:
        out.append("  <material_cell_offsets>false</material_cell_offsets>")
    if st.get("entropy_mesh") is not None:      # (dimension, lower_left, upper_right): Shannon entropy
        dim, ll, ur = st["entropy_mesh"]
        out += ['  <mesh id="9001">', f"    <dimension>{' '.join(str(int(d)) for d in dim)}</dimension>",
                f"    <lower_left>{_fmt(ll)}</lower_left>", f"    <upper_right>{_fmt(ur)}</upper_right>", "  </mesh>",
                "  <entropy_mesh>9001</entropy_mesh>"]
    for k, v in (extra or {}).items():
        out.append(f"  <{k}>{v}</{k}>")
    return out


def tallies_elements(m) -> list:
    out = []
    for mesh in m.meshes:
        out += [f'  <mesh id="{mesh.id}">', f"    <dimension>{' '.join(str(int(d)) for d in mesh.dimension)}</dimension>",
                f"    <lower_left>{_fmt(mesh.lower_left)}</lower_left>",
                f"    <upper_right>{_fmt(mesh.upper_right)}</upper_right>", "  </mesh>"]
    for f in m.filters:
        if f.type == "mesh":
            bins = str(int(f.bins))
        elif f.type in ("energy", "energyout"):
            bins = _fmt(f.bins)
        else:
            bins = " ".join(str(int(b)) for b in f.bins)
        out += [f'  <filter id="{f.id}" type="{f.type}">', f"    <bins>{bins}</bins>", "  </filter>"]
    for t in m.tallies:
        out.append(f'  <tally id="{t.id}">')
        if t.filters:
            out.append(f"    <filters>{' '.join(str(int(i)) for i in t.filters)}</filters>")
        out.append(f"    <scores>{' '.join(t.scores)}</scores>")
        if getattr(t, "estimator", "tracklength") != "tracklength":
            out.append(f"    <estimator>{t.estimator}</estimator>")
        out.append("  </tally>")
    return out


def model_xml(m, cross_sections, particles, batches, inactive, box, only_fissionable, run_CE, **kw) -> str:
    ind = lambda rows: ["  " + r for r in rows]
    out = ["<?xml version='1.0' encoding='utf-8'?>", "<model>", "  <materials>"]
    out += ind(materials_elements(m, cross_sections)) + ["  </materials>", "  <geometry>"]
    out += ind(geometry_elements(m)) + ["  </geometry>", "  <settings>"]
    out += ind(settings_elements(m, particles, batches, inactive, box, only_fissionable, run_CE, **kw))
    out += ["  </settings>"]
    if m.tallies:
        out += ["  <tallies>"] + ind(tallies_elements(m)) + ["  </tallies>"]
    out.append("</model>")
    return "\n".join(out) + "\n"


# ---------------------------------------------------------------------------------------------
# data libraries
# ---------------------------------------------------------------------------------------------
def _write_function(group, name, fn):
    """Function1D.to_hdf5 (openmc/data/function.py): a dataset with a `type` attribute"""
    if fn.kind == "poly":
        d = group.create_dataset(name, data=np.asarray(fn.data, float))
        d.attrs["type"] = b"Polynomial"
    else:
        x, y = fn.data
        d = group.create_dataset(name, data=np.vstack([np.asarray(x, float), np.asarray(y, float)]))
        d.attrs["type"] = b"Tabulated1D"
        d.attrs["breakpoints"] = np.array([len(x)], np.int32)
        d.attrs["interpolation"] = np.array([2], np.int32)   # lin-lin
    return d


_ELAW = {1: "level", 2: "maxwell", 3: "evaporation", 4: "watt"}   # OMCGPU_ELAW_* (include/omcgpu.h:94-98)


def write_nuclide(path, entry):
    """one synthetic nuclide (openmc_b200.nuclide.Nuclide, or a list of them = the same nuclide
    at several temperatures) -> <path> in the layout of IncidentNeutron.export_to_hdf5
    (openmc/data/neutron.py:354-431): kTs/<T>K, energy/<T>K, reactions/reaction_<mt>/<T>K/xs;
    products and distributions are temperature independent (taken from the first table)"""
    tables = list(entry) if isinstance(entry, (list, tuple)) else [entry]
    nuc = tables[0]
    labels = ["%dK" % int(round(t.kT / K_BOLTZMANN)) for t in tables]
    with omh5.File(path, "w") as f:
        f.attrs["filetype"] = b"data_neutron"
        f.attrs["version"] = np.array([3, 0], np.int32)      # HDF5_VERSION, include/openmc/constants.h:25
        g = f.create_group(nuc.name)
        g.attrs["Z"] = np.int32(nuc.Z)
        g.attrs["A"] = np.int32(nuc.A)
        g.attrs["metastable"] = np.int32(0)
        g.attrs["atomic_weight_ratio"] = float(nuc.awr)
        for t, label in zip(tables, labels):
            g.create_dataset(f"kTs/{label}", data=float(t.kT))
            g.create_dataset(f"energy/{label}", data=np.asarray(t.energy, float))
        rxs = g.create_group("reactions")
        for ir, r in enumerate(nuc.reactions):
            rg = rxs.create_group("reaction_%03d" % r.mt)
            rg.attrs["mt"] = np.int32(r.mt)
            rg.attrs["label"] = ("MT%d" % r.mt).encode()
            rg.attrs["Q_value"] = float(r.q_value)
            rg.attrs["center_of_mass"] = np.int32(1 if r.scatter_in_cm else 0)
            rg.attrs["redundant"] = np.int32(0)
            for t, label in zip(tables, labels):
                rt = t.reactions[ir]
                assert rt.mt == r.mt
                d = rg.create_dataset(f"{label}/xs", data=np.asarray(rt.xs, float))
                d.attrs["threshold_idx"] = np.int32(rt.threshold)
            for k, p in enumerate(r.products):
                pg = rg.create_group(f"product_{k}")
                pg.attrs["particle"] = b"neutron"
                pg.attrs["emission_mode"] = (b"prompt", b"delayed", b"total")[p.emission]
                if p.emission == 1:
                    pg.attrs["decay_rate"] = float(p.decay_rate)
                _write_function(pg, "yield", p.yield_)
                pg.attrs["n_distribution"] = np.int32(1)
                dg = pg.create_group("distribution_0")
                dg.attrs["type"] = b"uncorrelated"          # no `angle` group: isotropic
                if p.elaw in _ELAW:
                    eg = dg.create_group("energy")
                    eg.attrs["type"] = _ELAW[p.elaw].encode()
                    if p.elaw == 1:
                        eg.attrs["threshold"] = float(p.par[0])
                        eg.attrs["mass_ratio"] = float(p.par[1])
                    elif p.elaw in (2, 3):
                        eg.attrs["u"] = float(p.par[0])
                        _write_function(eg, "theta", p.fns[0])
                    else:
                        eg.attrs["u"] = float(p.par[0])
                        _write_function(eg, "a", p.fns[0])
                        _write_function(eg, "b", p.fns[1])
        if nuc.total_nu is not None:
            _write_function(g.create_group("total_nu"), "yield", nuc.total_nu)


def write_ce_library(directory, library) -> str:
    """-> path of cross_sections.xml listing one file per nuclide"""
    os.makedirs(directory, exist_ok=True)
    rows = ["<?xml version='1.0' encoding='utf-8'?>", "<cross_sections>"]
    for name, nuc in library.items():
        fn = f"{name}.h5"
        target = os.path.join(directory, fn)
        if not os.path.exists(target):
            write_nuclide(target, nuc)
        rows.append(f'  <library materials="{name}" path="{fn}" type="neutron"/>')
    rows.append("</cross_sections>")
    xml = os.path.join(directory, "cross_sections.xml")
    with open(xml, "w") as f:
        f.write("\n".join(rows) + "\n")
    return xml


def write_mgxs_library(path, lib):
    """openmc_b200.mgxs.MGXSLibrary -> mgxs.h5 (MGXSLibrary.export_to_hdf5 / XSdata.to_hdf5)"""
    G = lib.G
    with omh5.File(path, "w") as f:
        f.attrs["filetype"] = b"mgxs"
        f.attrs["version"] = np.array([1, 0], np.int32)
        f.attrs["energy_groups"] = np.int32(G)
        f.attrs["delayed_groups"] = np.int32(lib.num_delayed_groups)
        f.attrs["group structure"] = np.asarray(lib.group_edges, float)
        for x in lib.xsdatas.values():
            g = f.create_group(x.name)
            if x.atomic_weight_ratio is not None:
                g.attrs["atomic_weight_ratio"] = float(x.atomic_weight_ratio)
            g.attrs["fissionable"] = np.int8(1 if x.fissionable else 0)
            g.attrs["representation"] = x.representation.encode()
            angle = x.representation == "angle"
            if angle:
                g.attrs["num_azimuthal"] = np.int32(x.num_azimuthal)
                g.attrs["num_polar"] = np.int32(x.num_polar)
            g.attrs["scatter_shape"] = b"[G][G'][Order]"
            g.attrs["scatter_format"] = x.scatter_format.encode()
            g.attrs["order"] = np.int32(x.order)
            Np, Na = (x.num_polar, x.num_azimuthal) if angle else (1, 1)
            for T in x.temperatures:
                label = "%dK" % int(round(T))
                g.create_dataset(f"kTs/{label}", data=float(T) * K_BOLTZMANN)
                xs = g.create_group(label)
                d = x.data[T]
                for key in ("total", "absorption", "inverse-velocity"):
                    if key in d:
                        xs.create_dataset(key, data=d[key])
                if x.fissionable:
                    for key in ("fission", "kappa-fission", "chi", "chi-prompt", "chi-delayed", "nu-fission",
                                "prompt-nu-fission", "delayed-nu-fission", "beta", "decay-rate"):
                        if key in d:
                            xs.create_dataset(key, data=d[key])
                sm = np.asarray(d["scatter_matrix"], float)
                if sm.ndim == (4 if angle else 2):     # [G][G'] without an order axis
                    sm = sm[..., None]
                sm = sm.reshape((Np, Na, G, G, sm.shape[-1]))
                mm = d.get("multiplicity_matrix")
                if mm is not None:
                    mm = np.asarray(mm, float).reshape((Np, Na, G, G))
                bounds = np.zeros((Np, Na, G, 2), np.int32)
                flat_s, flat_m = [], []
                for p in range(Np):
                    for a in range(Na):
                        for gi in range(G):
                            row = sm[p, a, gi, :, 0] if x.scatter_format == "legendre" else sm[p, a, gi].sum(axis=1)
                            nz = np.nonzero(row)[0]
                            if nz.size:
                                bounds[p, a, gi] = (nz[0], nz[-1])
                        for gi in range(G):
                            lo, hi = bounds[p, a, gi]
                            flat_s.append(sm[p, a, gi, lo:hi + 1, :].reshape(-1))
                            if mm is not None:
                                flat_m.append(mm[p, a, gi, lo:hi + 1])
                sg = xs.create_group("scatter_data")
                sg.create_dataset("scatter_matrix", data=np.concatenate(flat_s))
                if mm is not None:
                    sg.create_dataset("multiplicity_matrix", data=np.concatenate(flat_m))
                bounds = bounds + 1
                if angle:
                    sg.create_dataset("g_min", data=bounds[..., 0])
                    sg.create_dataset("g_max", data=bounds[..., 1])
                else:
                    sg.create_dataset("g_min", data=bounds[0, 0, :, 0])
                    sg.create_dataset("g_max", data=bounds[0, 0, :, 1])


# ---------------------------------------------------------------------------------------------
# running the reference
# ---------------------------------------------------------------------------------------------
_RATE = re.compile(r"Calculation Rate \((\w+)\)\s+=\s+([0-9.eE+-]+) particles/second")
_TIME = re.compile(r"^\s*(Total time[^=]*|  Time in [^=]*|\s*Time [^=]*)=\s+([0-9.eE+-]+) seconds", re.M)
_KEFF = re.compile(r"Combined k-effective\s+=\s+([0-9.]+) \+/- ([0-9.]+)")


def run(directory, threads=None, event=False, max_in_flight=None, timeout=3600, binary=None):
    """execute the reference in `directory` (holds model.xml + data) -> dict"""
    cmd = [binary or BINARY]
    if threads:
        cmd += ["-s", str(int(threads))]
    if event:
        cmd.append("-e")
    env = dict(os.environ)
    env.pop("OPENMC_CROSS_SECTIONS", None)
    res = subprocess.run(cmd, cwd=directory, capture_output=True, text=True, timeout=timeout, env=env)
    out = res.stdout
    r = dict(returncode=res.returncode, stdout=out, stderr=res.stderr)
    if res.returncode != 0:
        return r
    m = _KEFF.search(out)
    if m:
        r["k_combined"] = (float(m.group(1)), float(m.group(2)))
    r["rates"] = {k: float(v) for k, v in _RATE.findall(out)}
    r["times"] = {k.strip(): float(v) for k, v in _TIME.findall(out)}
    sps = sorted(fn for fn in os.listdir(directory) if fn.startswith("statepoint.") and fn.endswith(".h5"))
    if sps:
        r["statepoint"] = omh5.File(os.path.join(directory, sps[-1]), "r")
    return r
