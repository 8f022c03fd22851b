"""Continuous-energy nuclide data: synthetic library + the load-time derivations of
data::nuclides.

ENDF data is not available offline (BASELINE.json north_star), so continuous-
energy runs use *synthetic* nuclides of stated grid size.  ``SyntheticNuclide``
generates, from a seed, the datasets an OpenMC HDF5 nuclide file holds
(docs/source/io_formats/nuclear_data.rst; writer openmc/data/neutron.py:354-431):
one temperature, an ascending energy grid, MT 2 / 18 / 51 / 102 pointwise cross
sections with threshold indices, polynomial nu-bar, Watt / Maxwell / level
secondary-energy laws, isotropic centre-of-mass elastic scattering -- the same
ingredients as the reference's own fake nuclide in
tests/unit_tests/test_temp_interp.py:20-86.

``build_ce_fields`` then restates what libopenmc derives at load time:

* Nuclide::create_derived  (xs_ table: total / absorption / fission / nu-fission)
                                                              src/nuclide.cpp:359-453
* data::energy_min/max, log_spacing                           src/simulation.cpp:682-761
* Nuclide::init_grid (log-hash table ``grid_index``)           src/nuclide.cpp:502-533
* Material::finalize / normalize_density                      src/material.cpp:392-467
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi as abi
from .model import MASS_NEUTRON, N_AVOGADRO
from .model import K_BOLTZMANN, Model, normalize_density

E_LOW, E_HIGH = 1.0e-5, 2.0e7


@dataclass
class Function1D:
    kind: str                 # 'poly' | 'tab'
    data: list                # coefficients, or (x[], y[])

    def __call__(self, x):
        if self.kind == "poly":
            y = 0.0
            for c in reversed(self.data):
                y = y * x + c
            return y
        xs, ys = self.data
        return float(np.interp(x, xs, ys))


@dataclass
class Product:
    emission: int             # 0 prompt, 1 delayed, 2 total
    yield_: Function1D
    elaw: int                 # OMCGPU_ELAW_*
    par: tuple = (0.0, 0.0)
    fns: tuple = (None, None)
    decay_rate: float = 0.0


@dataclass
class Reaction:
    mt: int
    threshold: int
    xs: np.ndarray
    q_value: float = 0.0
    scatter_in_cm: bool = True
    products: list = field(default_factory=list)


@dataclass
class Nuclide:
    name: str
    Z: int
    A: int
    awr: float
    kT: float
    energy: np.ndarray
    reactions: list           # MT-ordered: elastic first (src/nuclide.cpp:587-588)
    total_nu: Function1D | None = None

    @property
    def fissionable(self):
        return any(r.mt == 18 for r in self.reactions)


# per-nuclide knobs of the synthetic data: potential scattering [b], thermal
# (0.0253 eV) capture and fission [b], number of resonances and the scales of the
# capture / fission / elastic resonance peaks
PARAMS = {
    #          pot   cap_th  fis_th  n_res  rc     rf    re
    "U235":  (11.0,  99.0,  585.0,  250,   1.5,   6.0,  3.0),
    "U238":  (9.3,    2.7,    0.0,  200,   1.2,   0.0,  8.0),
    "Pu239": (10.0, 270.0,  750.0,  250,   6.0,  14.0,  3.0),
    "O16":   (3.8,  2.0e-4,   0.0,   10,   0.0,   0.0,  2.0),
    "He4":   (0.8,    0.0,    0.0,    4,   0.0,   0.0,  0.2),
    "Zr90":  (6.2,   0.011,   0.0,   60,   0.05,  0.0,  2.0),
    "Zr91":  (6.4,    1.2,    0.0,   80,   0.3,   0.0,  2.0),
    "Sn120": (4.7,   0.14,    0.0,   60,   0.2,   0.0,  2.0),
    "Fe56":  (11.4,   2.6,    0.0,   40,   0.2,   0.0,  4.0),
    "H1":    (20.4,  0.332,   0.0,    0,   0.0,   0.0,  0.0),
    "B10":   (2.1, 3840.0,    0.0,    0,   0.0,   0.0,  0.0),
}


def _params(name, Z, A, rng):
    if name in PARAMS:
        return PARAMS[name]
    # generic fission-product / actinide-like nuclide (used by the 300-nuclide stress case)
    heavy = A > 50
    return (float(rng.uniform(3.0, 12.0)), float(np.exp(rng.uniform(np.log(0.01), np.log(50.0)))),
            0.0, 120 if heavy else 10, float(rng.uniform(0.05, 2.0)) if heavy else 0.0, 0.0,
            float(rng.uniform(0.5, 4.0)))


def synthetic_nuclide(name, Z, A, n_grid=100000, fissionable=False, inelastic=False,
                      delayed=False, temperature=294.0, n2n=False) -> Nuclide:
    """Seeded synthetic nuclide: sigma(E) = smooth 1/v + sum of Lorentzian resonances."""
    rng = np.random.default_rng(1000 + Z * 1000 + A)
    awr = A * 0.99
    heavy = A > 50
    sigma_pot, cap_th, fis_th, n_res, rc, rf, re_ = _params(name, Z, A, rng)
    n_res = max(int(n_res), 1)
    lo, hi = (1.0, 3.0e4) if heavy else (1.0e4, 5.0e6)
    E_r = np.sort(np.exp(rng.uniform(np.log(lo), np.log(hi), n_res)))
    gam = E_r * rng.uniform(2e-4, 3e-3, n_res) + 0.03
    # crude Doppler broadening: resonances widen (and their peaks drop) with sqrt(T / 294 K); the
    # grid clusters follow the widths, so every temperature has its own energy grid
    broaden = math.sqrt(temperature / 294.0)
    gam = gam * broaden
    # grid: 60 % log-uniform + 40 % clustered on the resonances, plus exact end points
    n_smooth = int(0.6 * n_grid)
    g_smooth = np.exp(np.linspace(np.log(E_LOW), np.log(E_HIGH), n_smooth))
    n_cl = n_grid - n_smooth
    which = rng.integers(0, n_res, n_cl)
    g_cl = E_r[which] + gam[which] * rng.standard_cauchy(n_cl) * 0.7
    g_cl = np.clip(g_cl, E_LOW * 1.0001, E_HIGH * 0.9999)
    energy = np.unique(np.concatenate([g_smooth, g_cl]))
    while energy.size < n_grid:   # top up (duplicates removed by unique)
        extra = np.exp(rng.uniform(np.log(E_LOW), np.log(E_HIGH), n_grid - energy.size))
        energy = np.unique(np.concatenate([energy, extra]))
    energy = energy[:n_grid] if energy.size > n_grid else energy
    energy[0], energy[-1] = E_LOW, E_HIGH
    energy = np.sort(energy)

    def resonances(scale):
        out = np.zeros_like(energy)
        if scale == 0.0:
            return out
        peaks = scale * rng.lognormal(0.0, 0.8, n_res) * 100.0 / np.sqrt(np.maximum(E_r, 1.0)) / broaden
        for k in range(n_res):
            i0 = np.searchsorted(energy, E_r[k] - 400 * gam[k])
            i1 = np.searchsorted(energy, E_r[k] + 400 * gam[k])
            x = (energy[i0:i1] - E_r[k]) / (0.5 * gam[k])
            out[i0:i1] += peaks[k] / (1.0 + x * x)
        return out

    inv_v = np.sqrt(0.0253 / energy)
    elastic = sigma_pot / (1.0 + energy / 4.0e6) + resonances(re_)
    capture = cap_th * inv_v + resonances(rc)
    rx = [Reaction(2, 0, elastic, 0.0, True,
                   [Product(0, Function1D("poly", [1.0]), abi.OMCGPU_ELAW_NONE)])]
    total_nu = None
    if fissionable:
        fast = 0.6 * (energy > 0.8e6) * (1.0 - np.exp(-np.maximum(energy - 0.8e6, 0.0) / 6.0e5))
        fis = fis_th * inv_v + resonances(rf) + fast
        watt_a = Function1D("tab", ([E_LOW, E_HIGH], [0.988e6, 0.988e6 * 1.05]))
        watt_b = Function1D("tab", ([E_LOW, E_HIGH], [2.249e-6, 2.249e-6]))
        nu_p = Function1D("poly", [2.42, 1.4e-7])
        prods = [Product(0, nu_p, abi.OMCGPU_ELAW_WATT, (-2.0e7, 0.0), (watt_a, watt_b))]
        if delayed:
            for lam, y in ((0.0133, 0.0006), (0.0327, 0.0035), (0.3, 0.0026)):
                theta = Function1D("tab", ([E_LOW, E_HIGH], [0.4e6, 0.4e6]))
                prods.append(Product(1, Function1D("poly", [y]), abi.OMCGPU_ELAW_MAXWELL,
                                     (-2.0e7, 0.0), (theta, None), lam))
            total_nu = Function1D("poly", [2.42 + 0.0067, 1.4e-7])
        rx.append(Reaction(18, 0, fis, 193.0e6, False, prods))
    if inelastic:
        e_thr = 4.5e4 * (1.0 + 1.0 / awr)
        thr = int(np.searchsorted(energy, e_thr))
        e = energy[thr:]
        xs = 1.5 * (1.0 - np.exp(-(e - energy[thr]) / 3.0e5)) / (1.0 + e / 6.0e6)
        q = -4.5e4
        mass_ratio = (awr / (awr + 1.0)) ** 2
        rx.append(Reaction(51, thr, xs, q, True,
                           [Product(0, Function1D("poly", [1.0]), abi.OMCGPU_ELAW_LEVEL,
                                    ((awr + 1.0) / awr * abs(q), mass_ratio))]))
    if n2n:
        # (n,2n): integer yield 2 -> one secondary neutron per reaction (src/physics.cpp:1158-1162),
        # Maxwellian secondaries restricted to E_out <= E - U
        e_thr = 2.0e6 * (1.0 + 1.0 / awr)
        thr = int(np.searchsorted(energy, e_thr))
        e = energy[thr:]
        xs = 0.9 * (1.0 - np.exp(-(e - energy[thr]) / 1.5e6))
        theta = Function1D("tab", ([E_LOW, E_HIGH], [0.35e6, 0.35e6]))
        rx.append(Reaction(16, thr, xs, -1.9e6, False,
                           [Product(0, Function1D("poly", [2.0]), abi.OMCGPU_ELAW_MAXWELL,
                                    (1.9e6, 0.0), (theta, None))]))
    rx.append(Reaction(102, 0, capture, 5.0e6, False, []))
    rx.sort(key=lambda r: r.mt)   # HDF5 group-name order: reaction_002, _018, _051, _102
    return Nuclide(name, Z, A, awr, temperature * K_BOLTZMANN, energy, rx, total_nu)


def synthetic_library(names_ZA, n_grid=100000, temperatures=None, **kw):
    """names_ZA: [(name, Z, A, fissionable, inelastic, delayed)].  With `temperatures` (K) every
    entry is a LIST of tables, one per temperature in ascending order (an OpenMC nuclide file with
    several kTs / energy / reaction_xxx/<T>K groups)."""
    if temperatures:
        return {n: [synthetic_nuclide(n, Z, A, n_grid, f, i, d, temperature=T) for T in sorted(temperatures)]
                for n, Z, A, f, i, d in names_ZA}
    return {n: synthetic_nuclide(n, Z, A, n_grid, f, i, d, **kw) for n, Z, A, f, i, d in names_ZA}


def tables_of(entry):
    """library entry -> list of per-temperature tables (ascending kT)"""
    return list(entry) if isinstance(entry, (list, tuple)) else [entry]


def nuclide_temperatures(model: Model, names):
    """temperatures [K] every nuclide is needed at: get_temperatures, src/cross_sections.cpp --
    every material cell contributes its temperature(s) to the nuclides of its material, in cell order"""
    need = {n: [] for n in names}
    for c in model.cells:
        if c.fill_type != "material" or c.fill is None:
            continue
        T = c.temperature
        if T is None and c.fill.temperature is not None:
            T = c.fill.temperature
        if T is None:
            T = 293.6
        sq = math.sqrt(K_BOLTZMANN * T)
        T = sq * sq / K_BOLTZMANN                       # as the reference recovers it from sqrtkT_
        for n, _ in c.fill.components:
            if T not in need[n]:
                need[n].append(T)
    return need


def temperatures_to_read(tables, desired, tolerance):
    """Nuclide::Nuclide, TemperatureMethod::NEAREST (src/nuclide.cpp:84-165): indices of the
    library tables loaded for the desired temperatures, ascending"""
    avail = [float(round(t.kT / K_BOLTZMANN)) for t in tables]
    read = []
    for T in desired:
        best = min(range(len(avail)), key=lambda i: (abs(avail[i] - T), i))
        if abs(avail[best] - T) >= tolerance:
            raise ValueError(f"library does not contain cross sections at or near {T} K")
        if best not in read:
            read.append(best)
    return sorted(read, key=lambda i: avail[i])


PINCELL_NUCLIDES = [
    ("U235", 92, 235, True, True, True), ("U238", 92, 238, True, True, False),
    ("O16", 8, 16, False, False, False), ("He4", 2, 4, False, False, False),
    ("Zr90", 40, 90, False, True, False), ("Zr91", 40, 91, False, False, False),
    ("Sn120", 50, 120, False, False, False), ("Fe56", 26, 56, False, True, False),
    ("H1", 1, 1, False, False, False), ("B10", 5, 10, False, False, False),
]


# ------------------------------------------------------------------------------
# load-time derivations -> omcgpu_model fields
# ------------------------------------------------------------------------------
def _is_disappearance(mt):      # src/endf.cpp:60-72
    return (102 <= mt <= 117) or (600 <= mt <= 849) or mt in (155, 182, 191, 192, 193, 197)


def _is_inelastic_scatter(mt):  # src/endf.cpp:74-90
    if mt < 100:
        if mt in (18, 19, 20, 21, 38):
            return False
        return mt >= 5 and mt != 27
    if mt <= 200:
        return not _is_disappearance(mt)
    return 875 <= mt <= 891


def _log(x):
    return math.log(x)


def build_ce_fields(model: Model, library: dict) -> dict:
    st = model.settings
    # nuclide_map in order of first appearance over materials (src/material.cpp:236-243)
    names = []
    for m in model.materials:
        for n, _ in m.components:
            if n not in names:
                names.append(n)
    need = nuclide_temperatures(model, names)
    tol = float(st.get("temperature_tolerance", 10.0))
    nucs = []            # the TABLES (nuclide x loaded temperature), nuclide-major, ascending kT
    tab_ofs = [0]
    for n in names:
        tabs = tables_of(library[n])
        for k in temperatures_to_read(tabs, need[n], tol):
            nucs.append(tabs[k])
        tab_ofs.append(len(nucs))
    multi = len(nucs) != len(names)
    index = {n: i for i, n in enumerate(names)}       # mat_nuclide: nuclide index (== table index if not multi)
    M = int(st["n_log_bins"])
    # data::energy_min/max = intersection over nuclides (src/simulation.cpp:682-696)
    e_min = max(float(n.energy[0]) for n in nucs)
    e_max = min(float(n.energy[-1]) for n in nucs)
    spacing = _log(e_max / e_min) / M              # :757-760 and nuclide.cpp:510
    step = (M * spacing - 0.0) / max(1.0, float(M))
    umesh = [0.0 + step * float(k) for k in range(M + 1)]
    umesh[-1] = M * spacing                         # xt::linspace end point
    umesh = np.array(umesh)

    v = dict(n_nuclides=len(nucs), n_log_bins=M, log_spacing=spacing, energy_min=e_min,
             energy_max=e_max, n_nuclide_temps=len(names) if multi else 0,
             nuct_ofs=np.array(tab_ofs if multi else [0], np.int32))
    grid_ofs = [0]; energy = []; xs = []; gidx = []
    rx_ofs = [0]; fis_rx = []; n_prec = []; tot_nu = []
    rx_mt = []; rx_cm = []; rx_inel = []; rx_thr = []; rx_xs_ofs = [0]; rx_xs = []; rx_q = []
    rx_prod_ofs = [0]
    p_em = []; p_decay = []; p_yield = []; p_law = []; p_par = []; p_fn = []
    fn_type = []; fn_ofs = [0]; fn_data = []

    def add_fn(f: Function1D | None):
        if f is None:
            return -1
        if f.kind == "poly":
            fn_type.append(abi.OMCGPU_FN_POLYNOMIAL)
            fn_data.extend(float(c) for c in f.data)
        else:
            fn_type.append(abi.OMCGPU_FN_TABULATED1D)
            fn_data.extend(float(x) for x in f.data[0])
            fn_data.extend(float(y) for y in f.data[1])
        fn_ofs.append(len(fn_data))
        return len(fn_type) - 1

    for nuc in nucs:
        n = nuc.energy.size
        tab = np.zeros((n, 4))
        fissionable = False
        fr = -1
        for k, r in enumerate(nuc.reactions):       # create_derived, src/nuclide.cpp:359-431
            j, m_ = r.threshold, r.xs.size
            tab[j:j + m_, 0] += r.xs
            if _is_disappearance(r.mt):
                tab[j:j + m_, 1] += r.xs
            if r.mt in (18, 19, 20, 21, 38):
                fissionable = True
                tab[j:j + m_, 2] += r.xs
                tab[j:j + m_, 1] += r.xs
                if fr < 0:
                    fr = k
        if fissionable:                              # :444-453 nu(E, total) * fission
            nu_fn = nuc.total_nu if (nuc.total_nu is not None and st["create_delayed_neutrons"]) \
                else nuc.reactions[fr].products[0].yield_
            if nu_fn.kind == "poly":
                nu = np.zeros(n)
                for c in reversed(nu_fn.data):
                    nu = nu * nuc.energy + c
            else:
                nu = np.array([nu_fn(e) for e in nuc.energy])
            tab[:, 3] = nu * tab[:, 2]
        # init_grid, src/nuclide.cpp:502-533
        L = np.array([_log(e / e_min) for e in nuc.energy[1:]])
        gi = np.minimum(np.searchsorted(L, umesh, side="right"), n - 2).astype(np.int32)
        energy.append(nuc.energy); xs.append(tab); gidx.append(gi)
        grid_ofs.append(grid_ofs[-1] + n)
        fis_rx.append(fr)
        n_prec.append(sum(1 for p in nuc.reactions[fr].products if p.emission == 1) if fr >= 0 else 0)
        tot_nu.append(add_fn(nuc.total_nu))
        for r in nuc.reactions:
            rx_mt.append(r.mt); rx_cm.append(int(r.scatter_in_cm))
            rx_inel.append(int(_is_inelastic_scatter(r.mt))); rx_thr.append(r.threshold)
            rx_xs.append(r.xs); rx_xs_ofs.append(rx_xs_ofs[-1] + r.xs.size); rx_q.append(r.q_value)
            for p in r.products:
                p_em.append(p.emission); p_decay.append(p.decay_rate)
                p_yield.append(add_fn(p.yield_)); p_law.append(p.elaw)
                p_par.append(p.par); p_fn.append((add_fn(p.fns[0]), add_fn(p.fns[1])))
            rx_prod_ofs.append(len(p_em))
        rx_ofs.append(len(rx_mt))

    v.update(nuc_awr=np.array([n.awr for n in nucs]), nuc_kT=np.array([n.kT for n in nucs]),
             nuc_fissionable=np.array([int(f >= 0) for f in fis_rx], np.int32),
             nuc_grid_ofs=np.array(grid_ofs, np.int64), nuc_energy=np.concatenate(energy),
             nuc_xs=np.concatenate(xs).reshape(-1), nuc_grid_index=np.concatenate(gidx),
             nuc_rx_ofs=np.array(rx_ofs, np.int32), nuc_fission_rx=np.array(fis_rx, np.int32),
             nuc_n_precursor=np.array(n_prec, np.int32), nuc_total_nu_fn=np.array(tot_nu, np.int32),
             n_reactions=len(rx_mt), rx_mt=np.array(rx_mt, np.int32),
             rx_scatter_in_cm=np.array(rx_cm, np.int32), rx_inelastic=np.array(rx_inel, np.int32),
             rx_threshold=np.array(rx_thr, np.int32), rx_xs_ofs=np.array(rx_xs_ofs, np.int64),
             rx_xs=np.concatenate(rx_xs), rx_q_value=np.array(rx_q),
             rx_prod_ofs=np.array(rx_prod_ofs, np.int32), n_products=len(p_em),
             prod_emission=np.array(p_em or [0], np.int32), prod_decay_rate=np.array(p_decay or [0.0]),
             prod_yield_fn=np.array(p_yield or [0], np.int32), prod_elaw=np.array(p_law or [0], np.int32),
             prod_par=np.array(p_par or [(0.0, 0.0)]).reshape(-1),
             prod_fn=np.array(p_fn or [(-1, -1)], np.int32).reshape(-1),
             n_functions=len(fn_type), fn_type=np.array(fn_type or [0], np.int32),
             fn_ofs=np.array(fn_ofs, np.int32), fn_data=np.array(fn_data or [0.0]))

    # materials
    mofs = [0]; mnuc = []; mden = []; mfis = []
    gpcc = {}
    for m in model.materials:
        awr = [tables_of(library[n])[0].awr for n, _ in m.components]
        dens = normalize_density(m, awr)
        g = 0.0        # Material::density_gpcc_, src/material.cpp:458-466
        for d, a in zip(dens, awr):
            g += d * a * MASS_NEUTRON / N_AVOGADRO
        gpcc[m.id] = g
        mnuc += [index[n] for n, _ in m.components]
        mden += dens
        mofs.append(len(mnuc))
        mfis.append(int(any(tables_of(library[n])[0].fissionable for n, _ in m.components)))
    v.update(n_materials=len(model.materials),
             mat_id=np.array([m.id for m in model.materials], np.int32),
             mat_fissionable=np.array(mfis, np.int32), mat_nuc_ofs=np.array(mofs, np.int32),
             mat_nuclide=np.array(mnuc, np.int32), mat_atom_density=np.array(mden))
    # cell densities -> multipliers (finalize_cell_densities, src/geometry_aux.cpp:197-211)
    dofs = [0]; dmult = []
    for c in model.cells:
        if getattr(c, "density", None) is None:
            dmult.append(1.0)
        else:
            if c.fill_type != "material" or c.fill is None:
                raise ValueError(f"cell {c.id}: a density needs a (non-void) material")
            rho = np.atleast_1d(np.asarray(c.density, float))
            dmult += [float(r) / gpcc[c.fill.id] for r in rho]
        dofs.append(len(dmult))
    v.update(cell_density_ofs=np.array(dofs, np.int32), cell_density_mult=np.array(dmult or [1.0]))
    return v


def build_ce_model(model: Model, library: dict) -> abi.ModelHandle:
    model.settings["run_CE"] = 1
    v = {}
    v.update(model.settings_fields())
    v.update(model.flatten_geometry())
    v.update(build_ce_fields(model, library))
    v.update(model.flatten_tallies())
    v.update(mg_n_groups=0, mg_n_delayed=0)
    return abi.ModelHandle(v)
