"""Multigroup cross-section library: the load-time derivations of data::mg.

The device engine consumes the *host-loaded* tables of ``data::mg.macro_xs_``
(SURVEY.md Appendix B item 9, Appendix C).  Because the reference's HDF5 loader
is out of scope / unavailable here, this module restates -- operation for
operation, in scalar double arithmetic -- what the loader derives:

* what ``openmc.XSdata.to_hdf5`` writes (sparse g_min/g_max scatter layout)
                                                  openmc/mgxs_library.py:1952-2147
* ``XsData::from_hdf5`` (chi normalisation, zero fix-ups)   src/xsdata.cpp:81-395
* ``XsData::scatter_from_hdf5``                              src/xsdata.cpp:397-506
* ``ScattData{Legendre,Histogram,Tabular}::init``            src/scattdata.cpp:22-62,230-285,
                                                             436-500, 636-727
* ``convert_legendre_to_tabular``                            src/scattdata.cpp:839-900
* ``ScattData::base_combine`` / ``XsData::combine`` / macro ``Mgxs``
                                                             src/scattdata.cpp:66-154,
                                                             src/xsdata.cpp:508-560,
                                                             src/mgxs.cpp:310-428,
                                                             src/mgxs_interface.cpp:118-152
The end-to-end pin is the reference's own goldens (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi as abi
from .model import K_BOLTZMANN, Material, Model, normalize_density

DEFAULT_NMU = 33                 # include/openmc/constants.h:271
MACROSCOPIC_AWR = -2.0


def linspace(start, stop, n):
    """xt::linspace (vendor/xtensor xbuilder.hpp:206-221,460-468)."""
    step = (stop - start) / max(1.0, float(n - 1))
    out = [start + step * float(i) for i in range(n)]
    if n > 1:
        out[-1] = stop
    return out


def evaluate_legendre(n, data, x):
    """src/math_functions.cpp:98-121"""
    pnx = [1.0] * (n + 1)
    if n >= 1:
        pnx[1] = x
    for l in range(1, n):
        pnx[l + 1] = ((2 * l + 1) * x * pnx[l] - l * pnx[l - 1]) / (l + 1)
    val = 0.0
    for l in range(n + 1):
        val += (l + 0.5) * data[l] * pnx[l]
    return val


# ----------------------------------------------------------------------------
# user-facing containers (same vocabulary as openmc.XSdata / openmc.MGXSLibrary)
# ----------------------------------------------------------------------------
class XSdata:
    def __init__(self, name, n_groups, temperatures=(294.0,), representation="isotropic",
                 num_polar=1, num_azimuthal=1, num_delayed_groups=0):
        self.name = name
        self.G = n_groups
        self.temperatures = list(temperatures)
        self.representation = representation
        self.num_polar = num_polar if representation == "angle" else 1
        self.num_azimuthal = num_azimuthal if representation == "angle" else 1
        self.num_delayed_groups = num_delayed_groups
        self.order = 0
        self.scatter_format = "legendre"
        self.atomic_weight_ratio = None
        self.fissionable = False
        self.data = {T: {} for T in self.temperatures}

    def _set(self, key, value, temperature):
        T = self.temperatures[0] if temperature is None else temperature
        self.data[T][key] = np.array(value, dtype=float)

    def set_total(self, v, temperature=None): self._set("total", v, temperature)
    def set_absorption(self, v, temperature=None): self._set("absorption", v, temperature)
    def set_scatter_matrix(self, v, temperature=None): self._set("scatter_matrix", v, temperature)
    def set_multiplicity_matrix(self, v, temperature=None): self._set("multiplicity_matrix", v, temperature)
    def set_inverse_velocity(self, v, temperature=None): self._set("inverse-velocity", v, temperature)
    def set_decay_rate(self, v, temperature=None): self._set("decay-rate", v, temperature)
    def set_beta(self, v, temperature=None): self._set("beta", v, temperature)
    def set_chi(self, v, temperature=None): self._set("chi", v, temperature)
    def set_chi_prompt(self, v, temperature=None): self._set("chi-prompt", v, temperature)
    def set_chi_delayed(self, v, temperature=None): self._set("chi-delayed", v, temperature)
    def set_kappa_fission(self, v, temperature=None): self._set("kappa-fission", v, temperature)

    def _fiss(self, key, v, temperature):
        self._set(key, v, temperature)
        if np.sum(np.asarray(v)) > 0.0:       # openmc/mgxs_library.py set_nu_fission etc.
            self.fissionable = True

    def set_fission(self, v, temperature=None): self._fiss("fission", v, temperature)
    def set_nu_fission(self, v, temperature=None): self._fiss("nu-fission", v, temperature)
    def set_prompt_nu_fission(self, v, temperature=None): self._fiss("prompt-nu-fission", v, temperature)
    def set_delayed_nu_fission(self, v, temperature=None): self._fiss("delayed-nu-fission", v, temperature)

    @property
    def n_ang(self):
        return self.num_polar * self.num_azimuthal

    # -- library-side conversions (openmc/mgxs_library.py; used by the mg_basic recipe) --
    def _clone(self, name=None):
        import copy
        x = copy.deepcopy(self)
        if name is not None:
            x.name = name
        return x

    def convert_scatter_format(self, target_format, target_order):
        """XSdata.convert_scatter_format, openmc/mgxs_library.py:1787-1866, for a
        Legendre source: 'tabular' evaluates the series on linspace(-1, 1, n);
        'histogram' Simpson-integrates a 257-point sampling of every bin."""
        from scipy.integrate import simpson
        from scipy.special import eval_legendre
        assert self.scatter_format == "legendre" and target_format in ("tabular", "histogram")
        x = self._clone()
        x.scatter_format = target_format
        x.order = target_order
        n_src = self.order + 1
        for T in self.temperatures:
            orig = np.asarray(self.data[T]["scatter_matrix"], dtype=float)
            new = np.zeros(orig.shape[:-1] + (target_order,))
            if target_format == "tabular":
                mu = np.linspace(-1, 1, target_order)
                for imu in range(len(mu)):
                    for l in range(n_src):
                        new[..., imu] += ((l + 0.5) * eval_legendre(l, mu[imu]) * orig[..., l])
            else:
                mu = np.linspace(-1, 1, target_order + 1)
                for h in range(target_order):
                    mu_fine = np.linspace(mu[h], mu[h + 1], 257)     # _NMU, mgxs_library.py:34
                    fine = np.zeros(new.shape[:-1] + (257,))
                    for imu in range(len(mu_fine)):
                        for l in range(n_src):
                            fine[..., imu] += ((l + 0.5) * eval_legendre(l, mu_fine[imu]) * orig[..., l])
                    new[..., h] = simpson(fine, x=mu_fine)
            x.data[T]["scatter_matrix"] = new
        return x

    def convert_representation(self, target, num_polar, num_azimuthal):
        """isotropic -> angle: every array is copied into each (polar, azimuthal)
        bin, openmc/mgxs_library.py:1688-1785."""
        assert self.representation == "isotropic" and target == "angle"
        x = self._clone()
        x.representation = "angle"
        x.num_polar, x.num_azimuthal = num_polar, num_azimuthal
        for T in self.temperatures:
            for key, orig in self.data[T].items():
                x.data[T][key] = np.resize(orig, (num_polar, num_azimuthal) + np.asarray(orig).shape)
        return x

    def order_data(self):
        return self.order + 1 if self.scatter_format == "legendre" else self.order


class MGXSLibrary:
    def __init__(self, group_edges, num_delayed_groups=0):
        self.group_edges = [float(e) for e in group_edges]     # ascending, eV
        self.num_delayed_groups = num_delayed_groups
        self.xsdatas = {}

    @property
    def G(self):
        return len(self.group_edges) - 1

    def add_xsdata(self, x: XSdata):
        self.xsdatas[x.name] = x


# ----------------------------------------------------------------------------
# ScattData restatement
# ----------------------------------------------------------------------------
class ScattData:
    """Fields as include/openmc/scattdata.h:40-45 (+ subclass members)."""

    def __init__(self, fmt):
        self.fmt = fmt          # 'legendre' | 'histogram' | 'tabular'
        self.gmin = []; self.gmax = []
        self.energy = []; self.mult = []; self.dist = []
        self.scattxs = []
        self.max_val = []       # legendre
        self.mu = []; self.dmu = 0.0; self.fmu = []   # histogram / tabular

    # -- base_init, src/scattdata.cpp:22-62
    def base_init(self, order, gmin, gmax, energy, mult):
        G = len(energy)
        self.gmin = list(gmin); self.gmax = list(gmax)
        self.energy = [list(e) for e in energy]
        self.mult = [list(m) for m in mult]
        self.dist = []
        for gin in range(G):
            for go in range(len(self.mult[gin])):
                if self.mult[gin][go] == 0.0:
                    self.mult[gin][go] = 1.0
            norm = 0.0
            for e in self.energy[gin]:
                norm += e
            if norm != 0.0:
                self.energy[gin] = [e / norm for e in self.energy[gin]]
            self.dist.append([[0.0] * order for _ in range(gmax[gin] - gmin[gin] + 1)])

    def get_order(self):
        if self.fmt == "legendre":
            return len(self.dist[0][0]) - 1
        return len(self.dist[0][0])

    # -- init per representation
    def init(self, gmin, gmax, mult, coeffs):
        G = len(coeffs)
        order = len(coeffs[0][0])
        matrix = [[list(c) for c in row] for row in coeffs]
        if self.fmt == "legendre":           # src/scattdata.cpp:230-285
            self.scattxs = [0.0] * G
            for gin in range(G):
                for i_gout in range(gmax[gin] - gmin[gin] + 1):
                    self.scattxs[gin] += matrix[gin][i_gout][0]
            in_energy = []
            for gin in range(G):
                row = []
                for i_gout in range(gmax[gin] - gmin[gin] + 1):
                    norm = matrix[gin][i_gout][0]
                    row.append(norm)
                    if norm != 0.0:
                        matrix[gin][i_gout] = [n / norm for n in matrix[gin][i_gout]]
                in_energy.append(row)
            self.base_init(order, gmin, gmax, in_energy, mult)
            self.max_val = []
            for gin in range(G):
                ng = self.gmax[gin] - self.gmin[gin] + 1
                for i_gout in range(ng):
                    self.dist[gin][i_gout] = list(matrix[gin][i_gout])
                self.max_val.append([0.0] * ng)
            self.update_max_val()
        elif self.fmt == "histogram":        # src/scattdata.cpp:436-500
            self.scattxs = [0.0] * G
            for gin in range(G):
                for i_gout in range(len(matrix[gin])):
                    acc = 0.0
                    for x in matrix[gin][i_gout]:
                        acc += x
                    self.scattxs[gin] += acc
            in_energy = []
            for gin in range(G):
                row = []
                for i_gout in range(gmax[gin] - gmin[gin] + 1):
                    norm = 0.0
                    for x in matrix[gin][i_gout]:
                        norm += x
                    row.append(norm)
                    if norm != 0.0:
                        matrix[gin][i_gout] = [n / norm for n in matrix[gin][i_gout]]
                in_energy.append(row)
            self.base_init(order, gmin, gmax, in_energy, mult)
            self.mu = linspace(-1.0, 1.0, order + 1)
            self.dmu = 2.0 / order
            self.fmu = []
            for gin in range(G):
                ng = self.gmax[gin] - self.gmin[gin] + 1
                frow = []
                for i_gout in range(ng):
                    f = list(matrix[gin][i_gout])
                    d = self.dist[gin][i_gout]
                    d[0] = self.dmu * matrix[gin][i_gout][0]
                    for imu in range(1, order):
                        d[imu] = self.dmu * matrix[gin][i_gout][imu] + d[imu - 1]
                    norm = d[order - 1]
                    if norm > 0.0:
                        for imu in range(order):
                            f[imu] /= norm
                            d[imu] /= norm
                    frow.append(f)
                self.fmu.append(frow)
        else:                                # tabular, src/scattdata.cpp:636-727
            self.mu = linspace(-1.0, 1.0, order)
            self.dmu = 2.0 / (order - 1)
            dmu = self.dmu
            self.scattxs = [0.0] * G
            for gin in range(G):
                for i_gout in range(len(matrix[gin])):
                    for imu in range(1, order):
                        self.scattxs[gin] += 0.5 * dmu * (matrix[gin][i_gout][imu - 1] +
                                                          matrix[gin][i_gout][imu])
            in_energy = []
            for gin in range(G):
                row = []
                for i_gout in range(gmax[gin] - gmin[gin] + 1):
                    norm = 0.0
                    for imu in range(1, order):
                        norm += 0.5 * dmu * (matrix[gin][i_gout][imu - 1] + matrix[gin][i_gout][imu])
                    row.append(norm)
                in_energy.append(row)
            self.base_init(order, gmin, gmax, in_energy, mult)
            self.fmu = []
            for gin in range(G):
                ng = self.gmax[gin] - self.gmin[gin] + 1
                frow = []
                for i_gout in range(ng):
                    f = [max(x, 0.0) if x < 0.0 else x for x in matrix[gin][i_gout]]
                    d = self.dist[gin][i_gout]
                    norm = 0.0
                    for imu in range(1, order):
                        norm += 0.5 * dmu * (f[imu - 1] + f[imu])
                        d[imu] = norm
                    if norm > 0.0:
                        for imu in range(order):
                            f[imu] /= norm
                            d[imu] /= norm
                    frow.append(f)
                self.fmu.append(frow)

    def update_max_val(self):                # src/scattdata.cpp:289-325 (quirks kept)
        Nmu = 1001
        dmu = 2.0 / (Nmu - 1)
        for gin in range(len(self.max_val)):
            for i_gout in range(self.gmax[gin] - self.gmin[gin] + 1):
                d = self.dist[gin][i_gout]
                n = len(d) - 1
                mv = self.max_val[gin][i_gout]
                for imu in range(Nmu):
                    if imu == 0:
                        mu = -1.0
                    elif imu == Nmu - 1:
                        mu = 1.0
                    else:
                        mu = -1.0 + (imu - 1) * dmu
                    f = evaluate_legendre(n, d, mu)
                    if f > mv:
                        mv = f
                self.max_val[gin][i_gout] = mv * 1.1

    # -- get_xs for NU_SCATTER / SCATTER with a gout, src/scattdata.cpp:174-225
    def xs_nu_scatter(self, gin, gout):
        if gout < self.gmin[gin] or gout > self.gmax[gin]:
            return 0.0
        return self.scattxs[gin] * self.energy[gin][gout - self.gmin[gin]]

    def xs_scatter(self, gin, gout):
        if gout < self.gmin[gin] or gout > self.gmax[gin]:
            return 0.0
        i = gout - self.gmin[gin]
        return self.scattxs[gin] * (self.energy[gin][i] / self.mult[gin][i])

    def get_matrix(self, max_order):         # :407-432, 562-582, 779-799
        G = len(self.energy)
        if self.fmt == "legendre":
            order_dim = max_order + 1
            src = self.dist
        else:
            order_dim = self.get_order()
            src = self.fmu
        mat = [[[0.0] * order_dim for _ in range(G)] for _ in range(G)]
        for gin in range(G):
            for i_gout in range(len(self.energy[gin])):
                gout = i_gout + self.gmin[gin]
                for l in range(order_dim):
                    v = src[gin][i_gout][l] if l < len(src[gin][i_gout]) else 0.0
                    mat[gin][gout][l] = self.scattxs[gin] * self.energy[gin][i_gout] * v
        return mat

    def combine(self, those, scalars):       # :371-403, 586-624, 803-841 + base_combine :66-154
        if self.fmt == "legendre":
            max_order = 0
            for t in those:
                max_order = max(max_order, t.get_order())
            order_dim = max_order + 1
        else:
            max_order = those[0].get_order()
            order_dim = max_order
        G = len(those[0].energy)
        nusc = [[[0.0] * order_dim for _ in range(G)] for _ in range(G)]
        nu_p0 = [[0.0] * G for _ in range(G)]
        sc_p0 = [[0.0] * G for _ in range(G)]
        for i, that in enumerate(those):
            tm = that.get_matrix(max_order)
            for gin in range(G):
                for go in range(G):
                    for l in range(order_dim):
                        nusc[gin][go][l] += scalars[i] * tm[gin][go][l]
            for gin in range(G):
                for go in range(G):
                    nu_p0[gin][go] += scalars[i] * that.xs_nu_scatter(gin, go)
                    sc_p0[gin][go] += scalars[i] * that.xs_scatter(gin, go)
        mult = [[0.0] * G for _ in range(G)]
        for gin in range(G):
            for go in range(G):
                a, b = nu_p0[gin][go], sc_p0[gin][go]
                if b == 0.0:
                    q = 0.0 if (a == 0.0 or a != a) else math.copysign(1.7976931348623157e308, a)
                else:
                    q = a / b
                if q != q:
                    q = 0.0
                mult[gin][go] = q
        gmin = [0] * G; gmax = [0] * G
        sparse_scatter = []; sparse_mult = []
        for gin in range(G):
            g0 = 0
            while g0 < G and not any(nusc[gin][g0][l] != 0.0 for l in range(order_dim)):
                g0 += 1
            g1 = G - 1
            while g1 >= 0 and not any(nusc[gin][g1][l] != 0.0 for l in range(order_dim)):
                g1 -= 1
            if g0 > g1:
                g0 = gin; g1 = gin
            gmin[gin] = g0; gmax[gin] = g1
            sparse_scatter.append([list(nusc[gin][go]) for go in range(g0, g1 + 1)])
            sparse_mult.append([mult[gin][go] for go in range(g0, g1 + 1)])
        self.init(gmin, gmax, sparse_mult, sparse_scatter)


def convert_legendre_to_tabular(leg: ScattData, n_mu_setting=None) -> ScattData:
    """src/scattdata.cpp:839-900"""
    n_mu = n_mu_setting
    if n_mu is None:
        n_mu = 2 if leg.get_order() == 0 else DEFAULT_NMU
    tab = ScattData("tabular")
    tab.base_init(n_mu, leg.gmin, leg.gmax, leg.energy, leg.mult)
    tab.scattxs = list(leg.scattxs)
    tab.mu = linspace(-1.0, 1.0, n_mu)
    tab.dmu = 2.0 / (n_mu - 1)
    G = len(tab.energy)
    tab.fmu = []
    for gin in range(G):
        frow = []
        for i_gout in range(tab.gmax[gin] - tab.gmin[gin] + 1):
            c = leg.dist[gin][i_gout]
            f = [evaluate_legendre(len(c) - 1, c, tab.mu[imu]) for imu in range(n_mu)]
            f = [0.0 if x < 0.0 else x for x in f]
            d = tab.dist[gin][i_gout]
            norm = 0.0
            d[0] = 0.0
            for imu in range(1, n_mu):
                norm += 0.5 * tab.dmu * (f[imu - 1] + f[imu])
                d[imu] = norm
            if norm > 0.0:
                for imu in range(n_mu):
                    f[imu] /= norm
                    d[imu] /= norm
            frow.append(f)
        tab.fmu.append(frow)
    return tab


# ----------------------------------------------------------------------------
# XsData restatement (one temperature)
# ----------------------------------------------------------------------------
class XsDataT:
    def __init__(self, fissionable, fmt, n_ang, G, DG):
        self.fissionable = fissionable
        self.fmt = fmt; self.n_ang = n_ang; self.G = G; self.DG = DG
        z2 = lambda: [[0.0] * G for _ in range(n_ang)]
        self.total = z2(); self.absorption = z2(); self.inverse_velocity = z2()
        self.decay_rate = [[0.0] * DG for _ in range(n_ang)]
        if fissionable:
            self.fission = z2(); self.nu_fission = z2(); self.prompt_nu_fission = z2()
            self.kappa_fission = z2()
            self.delayed_nu_fission = [[[0.0] * G for _ in range(DG)] for _ in range(n_ang)]
            self.chi_prompt = [[[0.0] * G for _ in range(G)] for _ in range(n_ang)]
            self.chi_delayed = [[[[0.0] * G for _ in range(G)] for _ in range(DG)] for _ in range(n_ang)]
        self.scatter = [ScattData(fmt) for _ in range(n_ang)]


def _ang(arr, n_ang, tail_shape):
    """reshape library data to [n_ang][...] lists of python floats"""
    a = np.asarray(arr, dtype=float).reshape((n_ang,) + tuple(tail_shape))
    return a.tolist()


def load_xsdata(x: XSdata, T, final_fmt, settings, DG_lib) -> XsDataT:
    """XsData::from_hdf5 for the dataset openmc.XSdata.to_hdf5 would have written."""
    G, n_ang, DG = x.G, x.n_ang, DG_lib
    d = x.data[T]
    out = XsDataT(x.fissionable, final_fmt, n_ang, G, DG)
    if x.fissionable:                        # fission_from_hdf5 src/xsdata.cpp:353-392
        if "fission" in d:
            out.fission = _ang(d["fission"], n_ang, (G,))
        if "kappa-fission" in d:
            out.kappa_fission = _ang(d["kappa-fission"], n_ang, (G,))
        has_chi = "chi" in d or "chi-prompt" in d
        if has_chi and DG == 0:              # fission_vector_no_delayed :207-224
            chi = _ang(d["chi"], n_ang, (G,))
            for a in range(n_ang):
                s = 0.0
                for c in chi[a]:
                    s += c
                chi[a] = [c / s for c in chi[a]]
                for gin in range(G):
                    out.chi_prompt[a][gin] = list(chi[a])
            out.prompt_nu_fission = _ang(d["nu-fission"], n_ang, (G,))
        elif (not has_chi) and DG == 0:      # fission_matrix_no_delayed :335-351
            mat = _ang(d["nu-fission"], n_ang, (G, G))
            for a in range(n_ang):
                for gin in range(G):
                    s = 0.0
                    for go in range(G):
                        s += mat[a][gin][go]
                    out.prompt_nu_fission[a][gin] = s
                    out.chi_prompt[a][gin] = [mat[a][gin][go] / s for go in range(G)]
        else:
            raise NotImplementedError("delayed-group MGXS fission data: see mgxs_delayed.py")
        if DG == 0:
            out.nu_fission = [list(r) for r in out.prompt_nu_fission]
    if "decay-rate" in d:
        out.decay_rate = _ang(d["decay-rate"], n_ang, (DG,))
    out.absorption = _ang(d["absorption"], n_ang, (G,))
    if "inverse-velocity" in d:
        out.inverse_velocity = _ang(d["inverse-velocity"], n_ang, (G,))

    # ---- scatter: what to_hdf5 writes, then scatter_from_hdf5 :397-506
    order_data = x.order_data()
    sm = np.asarray(d["scatter_matrix"], dtype=float).reshape((n_ang, G, G, order_data))
    mm = d.get("multiplicity_matrix")
    if mm is not None:
        mm = np.asarray(mm, dtype=float).reshape((n_ang, G, G))
    if x.scatter_format == "legendre":
        max_order = settings.get("max_order")
        order_dim = (min(order_data - 1, max_order) if max_order is not None else order_data - 1) + 1
    else:
        order_dim = order_data
    for a in range(n_ang):
        gmin = []; gmax = []; coeffs = []; mult = []
        for gin in range(G):
            if x.scatter_format == "legendre":
                probe = sm[a, gin, :, 0]
            else:
                probe = np.sum(sm[a, gin, :, :], axis=1)
            nz = np.nonzero(probe)[0]
            g0, g1 = (0, 0) if len(nz) == 0 else (int(nz[0]), int(nz[-1]))
            gmin.append(g0); gmax.append(g1)
            coeffs.append([[float(sm[a, gin, go, l]) for l in range(order_dim)]
                           for go in range(g0, g1 + 1)])
            if mm is not None:
                mult.append([float(mm[a, gin, go]) for go in range(g0, g1 + 1)])
            else:
                mult.append([1.0] * (g1 - g0 + 1))
        if x.scatter_format == "legendre" and final_fmt == "tabular":
            leg = ScattData("legendre")
            leg.init(gmin, gmax, mult, coeffs)
            out.scatter[a] = convert_legendre_to_tabular(leg, settings.get("legendre_to_tabular_points"))
        else:
            out.scatter[a] = ScattData(x.scatter_format)
            out.scatter[a].init(gmin, gmax, mult, coeffs)
    for a in range(n_ang):                   # :105, :108-119
        out.absorption[a] = [1.0e-10 if v == 0.0 else v for v in out.absorption[a]]
    if "total" in d:
        out.total = _ang(d["total"], n_ang, (G,))
    else:
        for a in range(n_ang):
            for gin in range(G):
                out.total[a][gin] = out.absorption[a][gin] + out.scatter[a].scattxs[gin]
    for a in range(n_ang):
        out.total[a] = [1.0e-10 if v == 0.0 else v for v in out.total[a]]
    return out


def combine_xsdata(fissionable, fmt, n_ang, G, DG, those, scalars) -> XsDataT:
    """XsData::combine, src/xsdata.cpp:508-560 (onto a zero-initialised object)."""
    me = XsDataT(fissionable, fmt, n_ang, G, DG)
    for i, that in enumerate(those):
        s = scalars[i]
        for a in range(n_ang):
            for g in range(G):
                me.total[a][g] += s * that.total[a][g]
                me.absorption[a][g] += s * that.absorption[a][g]
        if i == 0:
            me.inverse_velocity = [list(r) for r in that.inverse_velocity]
        if that.fissionable:
            for a in range(n_ang):
                sum_p = 0.0
                for g in range(G):
                    sum_p += that.prompt_nu_fission[a][g]
                for g in range(G):
                    me.nu_fission[a][g] += s * that.nu_fission[a][g]
                    me.prompt_nu_fission[a][g] += s * that.prompt_nu_fission[a][g]
                    me.kappa_fission[a][g] += s * that.kappa_fission[a][g]
                    me.fission[a][g] += s * that.fission[a][g]
                for gi in range(G):
                    for go in range(G):
                        me.chi_prompt[a][gi][go] += (s * sum_p) * that.chi_prompt[a][gi][go]
        for a in range(n_ang):
            for dg in range(DG):
                me.decay_rate[a][dg] += s * that.decay_rate[a][dg]
    if fissionable:
        for a in range(n_ang):
            for gi in range(G):
                tot = 0.0
                for go in range(G):
                    tot += me.chi_prompt[a][gi][go]
                me.chi_prompt[a][gi] = [c / tot if tot != 0.0 else float("nan")
                                        for c in me.chi_prompt[a][gi]]
    for a in range(n_ang):
        me.scatter[a] = ScattData(fmt)
        me.scatter[a].combine([t.scatter[a] for t in those], scalars)
    return me


# ----------------------------------------------------------------------------
# data::mg for a model
# ----------------------------------------------------------------------------
@dataclass
class MicroMgxs:
    name: str
    awr: float
    fissionable: bool
    fmt: str
    n_pol: int
    n_azi: int
    isotropic: bool
    kTs: list
    xs: list


def _nuclide_temperatures(model: Model, nuc_names, mat_nuclides):
    """get_temperatures (src/cross_sections.cpp): temperatures (K) of every cell
    containing each nuclide, in cell order, without duplicates."""
    temps = {n: [] for n in nuc_names}
    for c in model.cells:
        if c.fill_type != "material" or c.fill is None:
            continue
        T = c.temperature
        if T is None and c.fill.temperature is not None:
            T = c.fill.temperature
        if T is None:
            T = 293.6
        # the reader stores sqrt(kT) and converts back: T = sqrtkT^2 / k
        sq = math.sqrt(K_BOLTZMANN * T)
        Tc = sq * sq / K_BOLTZMANN
        for n in mat_nuclides[c.fill.id]:
            if Tc not in temps[n]:
                temps[n].append(Tc)
    return temps


def build_mg_fields(model: Model, lib: MGXSLibrary) -> dict:
    """Everything omcgpu_model needs for a multigroup run."""
    st = model.settings
    G, DG = lib.G, lib.num_delayed_groups
    # nuclide map in order of first appearance over materials (src/material.cpp:236-243)
    nuc_names = []
    mat_nuclides = {}
    for m in model.materials:
        names = [m.macroscopic] if m.macroscopic else [n for n, _ in m.components]
        mat_nuclides[m.id] = names
        for n in names:
            if n not in nuc_names:
                nuc_names.append(n)
    temps = _nuclide_temperatures(model, nuc_names, mat_nuclides)

    micros = {}
    for n in nuc_names:                      # Mgxs::Mgxs(hid_t...) src/mgxs.cpp:272-306
        x = lib.xsdatas[n]
        avail = sorted(round(T * K_BOLTZMANN / K_BOLTZMANN) for T in x.temperatures)
        to_read = []
        for T in temps[n]:                   # NEAREST, :96-125
            best = min(range(len(avail)), key=lambda i: (abs(avail[i] - T), i))
            t_act = avail[best]
            if abs(t_act - T) < st["temperature_tolerance"]:
                if round(t_act) not in to_read:
                    to_read.append(round(t_act))
            else:
                raise ValueError(f"MGXS library does not contain cross sections for {n} at or near {round(T)} K")
        to_read.sort()
        final_fmt = x.scatter_format
        if st["legendre_to_tabular"] and x.scatter_format == "legendre":
            final_fmt = "tabular"
        xs = []
        kTs = []
        for Tr in to_read:
            Tlib = [T for T in x.temperatures if int(round(T)) == Tr][0]
            kTs.append(Tlib * K_BOLTZMANN)
            xs.append(load_xsdata(x, Tlib, final_fmt, st, DG))
        awr = x.atomic_weight_ratio if x.atomic_weight_ratio is not None else MACROSCOPIC_AWR
        micros[n] = MicroMgxs(n, awr, x.fissionable, final_fmt, x.num_polar, x.num_azimuthal,
                              x.representation == "isotropic", kTs, xs)

    # materials: atom densities + kTs from cells (get_mat_kTs src/mgxs_interface.cpp:154-183)
    v = {}
    mat_kTs = {m.id: [] for m in model.materials}
    for c in model.cells:
        if c.fill_type != "material" or c.fill is None:
            continue
        T = c.temperature
        if T is None and c.fill.temperature is not None:
            T = c.fill.temperature
        if T is None:
            T = 293.6
        sq = math.sqrt(K_BOLTZMANN * T)
        kT = sq * sq
        if kT not in mat_kTs[c.fill.id]:
            mat_kTs[c.fill.id].append(kT)

    nuc_index = {n: i for i, n in enumerate(nuc_names)}
    mat_nuc_ofs = [0]; mat_nuc = []; mat_dens = []
    macros = []
    for m in model.materials:
        names = mat_nuclides[m.id]
        if m.macroscopic:
            comp_mat = Material(id=m.id, components=[(m.macroscopic, 1.0)], density=m.density,
                                units="macro")
        else:
            comp_mat = m
        dens = normalize_density(comp_mat, [micros[n].awr for n in names])
        mat_nuc += [nuc_index[n] for n in names]
        mat_dens += dens
        mat_nuc_ofs.append(len(mat_nuc))
        mic = [micros[n] for n in names]
        fissionable = any(mc.fissionable for mc in mic)
        blocks = []
        for kT in mat_kTs[m.id]:             # macro Mgxs ctor src/mgxs.cpp:310-414 (NEAREST)
            those = []
            for mc in mic:
                best = min(range(len(mc.kTs)), key=lambda i: (abs(mc.kTs[i] - kT), i))
                if abs(mc.kTs[best] - kT) >= K_BOLTZMANN * st["temperature_tolerance"]:
                    raise ValueError(f"MGXS Library does not contain cross section for {m.name} "
                                     f"at or near {round(kT / K_BOLTZMANN)} K.")
                those.append(mc.xs[best])
            blocks.append(combine_xsdata(fissionable, mic[0].fmt, mic[0].n_pol * mic[0].n_azi,
                                         G, DG, those, dens))
        macros.append((m, mic[0], fissionable, mat_kTs[m.id], blocks))

    v.update(n_materials=len(model.materials),
             mat_id=np.array([m.id for m in model.materials], np.int32),
             mat_fissionable=np.array([int(f) for _, _, f, _, _ in macros], np.int32),
             mat_nuc_ofs=np.array(mat_nuc_ofs, np.int32), mat_nuclide=np.array(mat_nuc, np.int32),
             mat_atom_density=np.array(mat_dens))

    # ---- flatten blocks
    energy_bins = list(reversed(lib.group_edges))          # descending (energy_bins_)
    bin_avg = [0.5 * (energy_bins[i] + energy_bins[i + 1]) for i in range(G)]
    v.update(mg_n_groups=G, mg_n_delayed=DG, mg_energy_bin_avg=np.array(bin_avg),
             mg_energy_bins=np.array(energy_bins),
             energy_min=energy_bins[-1], energy_max=energy_bins[0])
    n_temps = []; n_pol = []; n_azi = []; iso = []; fis = []; blk_ofs = [0]; kts = []
    ang_ofs = [0]
    cols = {k: [] for k in ("total", "absorption", "nu_fission", "prompt_nu_fission", "fission",
                            "inverse_velocity", "delayed_nu_fission", "chi_prompt", "chi_delayed",
                            "decay_rate")}
    sc = {k: [] for k in ("format", "order", "gmin", "gmax", "gofs", "dofs", "scattxs", "energy", "mult",
                          "maxval", "dist", "fmu", "mu_ofs", "mu", "dmu")}
    fmt_code = {"legendre": abi.OMCGPU_SCATT_LEGENDRE, "histogram": abi.OMCGPU_SCATT_HISTOGRAM,
                "tabular": abi.OMCGPU_SCATT_TABULAR}
    for m, m0, fissionable, kTs, blocks in macros:
        n_temps.append(len(blocks)); n_pol.append(m0.n_pol); n_azi.append(m0.n_azi)
        iso.append(int(m0.isotropic)); fis.append(int(fissionable))
        for kT, b in zip(kTs, blocks):
            kts.append(kT)
            for a in range(b.n_ang):
                cols["total"] += b.total[a]; cols["absorption"] += b.absorption[a]
                cols["inverse_velocity"] += b.inverse_velocity[a]
                cols["decay_rate"] += b.decay_rate[a]
                if fissionable:
                    cols["nu_fission"] += b.nu_fission[a]
                    cols["prompt_nu_fission"] += b.prompt_nu_fission[a]
                    cols["fission"] += b.fission[a]
                    for dg in range(DG):
                        cols["delayed_nu_fission"] += b.delayed_nu_fission[a][dg]
                    for gi in range(G):
                        cols["chi_prompt"] += b.chi_prompt[a][gi]
                    for dg in range(DG):
                        for gi in range(G):
                            cols["chi_delayed"] += b.chi_delayed[a][dg][gi]
                else:
                    cols["nu_fission"] += [0.0] * G; cols["prompt_nu_fission"] += [0.0] * G
                    cols["fission"] += [0.0] * G
                    cols["delayed_nu_fission"] += [0.0] * (DG * G)
                    cols["chi_prompt"] += [0.0] * (G * G)
                    cols["chi_delayed"] += [0.0] * (DG * G * G)
                s = b.scatter[a]
                order = len(s.dist[0][0])
                sc["format"].append(fmt_code[s.fmt]); sc["order"].append(order)
                sc["gmin"] += s.gmin; sc["gmax"] += s.gmax; sc["scattxs"] += s.scattxs
                sc["mu_ofs"].append(len(sc["mu"])); sc["mu"] += (s.mu if s.mu else [0.0])
                sc["dmu"].append(s.dmu)
                for gin in range(G):
                    sc["gofs"].append(len(sc["energy"]))
                    sc["dofs"].append(len(sc["dist"]))
                    for i_gout in range(len(s.energy[gin])):
                        sc["energy"].append(s.energy[gin][i_gout])
                        sc["mult"].append(s.mult[gin][i_gout])
                        sc["maxval"].append(s.max_val[gin][i_gout] if s.fmt == "legendre" else 0.0)
                        sc["dist"] += s.dist[gin][i_gout]
                        sc["fmu"] += (s.fmu[gin][i_gout] if s.fmt != "legendre" else [0.0] * order)
            ang_ofs.append(ang_ofs[-1] + b.n_ang)
        blk_ofs.append(len(kts))
    v.update(mgm_n_temps=np.array(n_temps, np.int32), mgm_n_pol=np.array(n_pol, np.int32),
             mgm_n_azi=np.array(n_azi, np.int32), mgm_isotropic=np.array(iso, np.int32),
             mgm_fissionable=np.array(fis, np.int32), mgm_blk_ofs=np.array(blk_ofs, np.int32),
             mgm_kTs=np.array(kts), blk_ang_ofs=np.array(ang_ofs, np.int32),
             mg_total=np.array(cols["total"]), mg_absorption=np.array(cols["absorption"]),
             mg_nu_fission=np.array(cols["nu_fission"]),
             mg_prompt_nu_fission=np.array(cols["prompt_nu_fission"]),
             mg_fission=np.array(cols["fission"]),
             mg_inverse_velocity=np.array(cols["inverse_velocity"]),
             mg_delayed_nu_fission=np.array(cols["delayed_nu_fission"] or [0.0]),
             mg_chi_prompt=np.array(cols["chi_prompt"]),
             mg_chi_delayed=np.array(cols["chi_delayed"] or [0.0]),
             mg_decay_rate=np.array(cols["decay_rate"] or [0.0]),
             sc_format=np.array(sc["format"], np.int32), sc_order=np.array(sc["order"], np.int32),
             sc_gmin=np.array(sc["gmin"], np.int32), sc_gmax=np.array(sc["gmax"], np.int32),
             sc_gofs=np.array(sc["gofs"], np.int32), sc_dofs=np.array(sc["dofs"], np.int32),
             sc_scattxs=np.array(sc["scattxs"]),
             sc_energy=np.array(sc["energy"]), sc_mult=np.array(sc["mult"]),
             sc_maxval=np.array(sc["maxval"]), sc_dist=np.array(sc["dist"]),
             sc_fmu=np.array(sc["fmu"]), sc_mu_ofs=np.array(sc["mu_ofs"], np.int32),
             sc_mu=np.array(sc["mu"]), sc_dmu=np.array(sc["dmu"]))
    return v


def build_mg_model(model: Model, lib: MGXSLibrary) -> abi.ModelHandle:
    model.settings["run_CE"] = 0
    v = {}
    v.update(model.settings_fields())
    v.update(model.flatten_geometry())
    v.update(build_mg_fields(model, lib))
    v.update(model.flatten_tallies(lib.G, lib.group_edges))
    # cell densities -> multipliers (finalize_cell_densities, src/geometry_aux.cpp:197-211; a macroscopic
    # material's density_gpcc_ is its one atom density with awr = 1, src/material.cpp:458-466)
    from .model import MASS_NEUTRON, N_AVOGADRO
    dofs = [0]; dmult = []
    for c in model.cells:
        if getattr(c, "density", None) is None:
            dmult.append(1.0)
        else:
            if c.fill_type != "material" or c.fill is None:
                raise ValueError(f"cell {c.id}: a density needs a (non-void) material")
            rho_m = 1.0 if c.fill.density is None else float(c.fill.density)
            gpcc = 0.0
            gpcc += rho_m * 1.0 * MASS_NEUTRON / N_AVOGADRO
            dmult += [float(r) / gpcc for r in np.atleast_1d(np.asarray(c.density, float))]
        dofs.append(len(dmult))
    v.update(cell_density_ofs=np.array(dofs, np.int32), cell_density_mult=np.array(dmult or [1.0]))
    # unused CE fields
    v.update(n_nuclides=0, n_log_bins=0, log_spacing=1.0, n_reactions=0, n_products=0, n_functions=0)
    return abi.ModelHandle(v)
