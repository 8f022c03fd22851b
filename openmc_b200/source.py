"""Initial fission source on the host (out of the hot path: sampled once per
run by libopenmc, src/source.cpp:601-625).  Restates IndependentSource::sample
(src/source.cpp:344-399) for the source the configs use -- uniform box,
isotropic, Watt(0.988 MeV, 2.249 /MeV) energy, t = 0 -- vectorised with numpy;
the 'fissionable' spatial constraint is evaluated with the device find_cell
probe (omcgpu_find_cell)."""
from __future__ import annotations

import numpy as np

from . import _abi as abi
from .rng import STREAM_SOURCE, future_seed_v, prn_v


def sample_initial_source(engine, n, box, only_fissionable=False, master_seed=1, stride=152917,
                          id_offset=0, watt=(0.988e6, 2.249e-6)):
    model = engine.model
    run_CE = bool(model["run_CE"])
    ids = np.arange(1, n + 1, dtype=np.uint64) + np.uint64(id_offset)
    with np.errstate(over="ignore"):
        seed = future_seed_v(ids * np.uint64(stride),
                             np.full(n, master_seed + STREAM_SOURCE, np.uint64))
    box = np.asarray(box, float)
    r = np.zeros((n, 3))
    todo = np.arange(n)
    fissionable = np.asarray(model["mat_fissionable"])
    while todo.size:
        s = seed[todo]
        xi = np.zeros((todo.size, 3))
        for k in range(3):
            xi[:, k], s = prn_v(s)
        seed[todo] = s
        cand = box[:3] + xi * (box[3:] - box[:3])
        cell, _, mat, _ = engine.find_cell(cand)
        ok = cell >= 0
        if only_fissionable:
            ok &= (mat >= 0) & (fissionable[np.clip(mat, 0, None)] != 0)
        r[todo[ok]] = cand[ok]
        todo = todo[~ok]
    xi, seed = prn_v(seed)
    phi = 0.0 + (2.0 * np.pi - 0.0) * xi
    xi, seed = prn_v(seed)
    mu = -1.0 + (1.0 - (-1.0)) * xi
    u = np.stack([mu, np.sqrt(1.0 - mu * mu) * np.cos(phi), np.sqrt(1.0 - mu * mu) * np.sin(phi)], 1)
    a, b = watt
    E = np.zeros(n)
    todo = np.arange(n)
    e_max = float(model["energy_max"])
    while todo.size:
        s = seed[todo]
        r1, s = prn_v(s); r2, s = prn_v(s); r3, s = prn_v(s)
        c = np.cos(np.pi / 2.0 * r3)
        w = -a * (np.log(r1) + np.log(r2) * c * c)
        xi, s = prn_v(s)
        e = w + 0.25 * a * a * b + (-1.0 + 2.0 * xi) * np.sqrt(a * a * b * w)
        seed[todo] = s
        ok = (e < e_max) & (e > 0.0)
        E[todo[ok]] = e[ok]
        todo = todo[~ok]
    bank = np.zeros(n, abi.site_dtype)
    bank["r"] = r
    bank["u"] = u
    bank["wgt"] = 1.0
    bank["parent_nuclide"] = -1
    if run_CE:
        bank["E"] = E
    else:
        G = int(model["mg_n_groups"])
        rev = np.asarray(model["mg_energy_bins"])[::-1]
        # lower_bound_index(rev, E) then G - g - 1 (src/source.cpp:651-656)
        lb = np.searchsorted(rev, E, side="left") - 1
        lb = np.where(E == rev[0], 0, lb)
        bank["E"] = (G - lb - 1).astype(float)
    return bank
