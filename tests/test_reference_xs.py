"""Continuous-energy lookup against the REFERENCE ITSELF: Nuclide::init_grid and
Nuclide::calculate_xs of src/nuclide.cpp, compiled unmodified into
oracle/_ref/libref_geom.so, on the synthetic pointwise tables.  The hash table, every grid
index (incl. energies ON grid points, one ulp either side, and on hash-bin edges) and the
density-weighted sums must be identical in the oracle; the CUDA path is held to the
committed reference vectors (indices bit-exact, cross sections <= 1e-12: north_star)."""
import os

import numpy as np
import pytest

import cases
import ref_geometry
from oracle.binding import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not os.path.exists(ref_geometry.LIB),
                               reason="oracle/_ref/libref_geom.so not built (no /root/reference here)")


@needs_ref
def test_hash_table_and_lookup_equal_reference():
    model, box, of, libx = cases.build_ce(n_grid=4000)
    ref = ref_geometry.ReferenceXS(model)
    gi = np.asarray(model["nuc_grid_index"]).reshape(ref.nn, ref.n_log_bins + 1)
    for k in range(ref.nn):
        assert np.array_equal(ref.grid_index(k), gi[k])        # Nuclide::init_grid
    E = ref.in_range(cases.xs_probe_energies(model, libx, 30000))
    o = Oracle(model)
    for mat in range(int(model["n_materials"])):
        n_nuc = int(model["mat_nuc_ofs"][mat + 1] - model["mat_nuc_ofs"][mat])
        if n_nuc == 0:
            continue
        xs_r, ig_r = ref.macro(mat, E)
        xs_o, ig_o = o.calculate_xs(E, np.full(E.size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_o[:, :n_nuc], ig_r)
        assert np.array_equal(xs_o, xs_r)


def _fixture():
    model = cases.load_model_npz(os.path.join(GOLDEN, "ce_model.npz"))
    return model, np.load(os.path.join(GOLDEN, "refxs_ce.npz"))


def test_committed_reference_vectors_are_reproduced_by_the_oracle():
    model, v = _fixture()
    o = Oracle(model)
    for mat in range(int(model["n_materials"])):
        if f"xs{mat}" not in v.files:
            continue
        n_nuc = v[f"ig{mat}"].shape[1]
        xs_o, ig_o = o.calculate_xs(v["E"], np.full(v["E"].size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_o[:, :n_nuc], v[f"ig{mat}"]) and np.array_equal(xs_o, v[f"xs{mat}"])


def test_committed_reference_collisions_are_reproduced_by_the_oracle():
    """tests/golden/refcoll_ce.npz (written by the reference's own collision(),
    make_reference_collisions.py) without the reference tree -- this one travels to the GPU box:
    the oracle leaves every neutron as the reference did, analog and with survival biasing"""
    model = cases.load_model_npz(os.path.join(GOLDEN, "ce_model.npz"))
    v = np.load(os.path.join(GOLDEN, "refcoll_ce.npz"))
    mat, E, u, seeds = (np.ascontiguousarray(v[k]) for k in ("material", "E_in", "u_in", "seed_in"))
    for tag, survival in (("a", 0), ("s", 1)):
        model.struct.survival_biasing = survival
        o = Oracle(model); o.load()
        q = ref_geometry.oracle_collisions(o, mat, E, u, seeds, float(v["k_eff"]), float(v[f"{tag}_wgt0"]))
        for name in ref_geometry.COLLISION_FIELDS:
            assert np.array_equal(q[name], v[f"{tag}_{name}"]), (tag, name)
        assert int(v[f"{tag}_site_ofs"][-1]) > 50


@pytest.mark.gpu
def test_gpu_lookup_equals_reference_vectors(gpu_engine):
    model, v = _fixture()
    gpu_engine.init(model, 0, 1024)
    for mat in range(int(model["n_materials"])):
        if f"xs{mat}" not in v.files:
            continue
        n_nuc = v[f"ig{mat}"].shape[1]
        xs_g, ig_g = gpu_engine.calculate_xs(v["E"], np.full(v["E"].size, mat, np.int32), n_nuc)
        assert np.array_equal(ig_g[:, :n_nuc], v[f"ig{mat}"])                  # bit-exact
        np.testing.assert_allclose(xs_g, v[f"xs{mat}"], rtol=1e-12, atol=0)     # north_star tolerance
    gpu_engine.finalize()


@needs_ref
def test_function1d_evaluation_equals_reference():
    """every Polynomial / Tabulated1D of the synthetic library (nu(E), Watt a(E) b(E), Maxwell
    theta(E), yields) through the reference's own operator() (src/endf.cpp, compiled
    unmodified): identical values inside, on and outside the tabulated range"""
    import ctypes as C
    from openmc_b200 import _abi as abi
    model, box, of, libx = cases.build_ce(n_grid=2000)
    L = C.CDLL(ref_geometry.LIB)
    o = Oracle(model)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn_ofs = np.asarray(model["fn_ofs"]); fn_type = np.asarray(model["fn_type"]); data = np.asarray(model["fn_data"])
    g = np.random.default_rng(3)
    n_tab = n_pol = 0
    for f in range(int(model["n_functions"])):
        d = np.ascontiguousarray(data[fn_ofs[f]:fn_ofs[f + 1]])
        x = np.exp(g.uniform(np.log(1e-6), np.log(3e7), 3000))
        want = np.zeros_like(x)
        if fn_type[f] == abi.OMCGPU_FN_POLYNOMIAL:
            assert L.ref_eval_polynomial(p(d), C.c_int(d.size), C.c_int64(x.size), p(x), p(want)) == 0
            n_pol += 1
        else:
            k = d.size // 2
            xs, ys = np.ascontiguousarray(d[:k]), np.ascontiguousarray(d[k:])
            x = np.ascontiguousarray(np.concatenate([x, xs, np.nextafter(xs, 0), np.nextafter(xs, np.inf),
                                                     [xs[0] * 0.5, xs[-1] * 2.0]]))
            want = np.zeros_like(x)
            assert L.ref_eval_tabulated(p(xs), p(ys), C.c_int(k), C.c_int64(x.size), p(x), p(want)) == 0
            n_tab += 1
        assert np.array_equal(o.eval_function(f, x), want), f
    assert n_tab > 0 and n_pol > 0


@needs_ref
def test_secondary_energy_laws_equal_reference():
    """every reaction product of the synthetic library (level inelastic, Maxwell, evaporation,
    Watt behind an isotropic UncorrelatedAngleEnergy) sampled by the reference's own
    distribution_energy.cpp from the same seeds: identical mu, outgoing energy and seed
    afterwards (the rejection loops consume the same random numbers)"""
    import ctypes as C
    from openmc_b200 import _abi as abi
    model, box, of, libx = cases.build_ce(n_grid=2000)
    L = C.CDLL(ref_geometry.LIB)
    L.ref_geom_error.restype = C.c_char_p
    o = Oracle(model)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn_ofs = np.asarray(model["fn_ofs"]); fn_type = np.asarray(model["fn_type"]); data = np.asarray(model["fn_data"])
    elaw = np.asarray(model["prod_elaw"]); par = np.asarray(model["prod_par"]).reshape(-1, 2)
    pfn = np.asarray(model["prod_fn"]).reshape(-1, 2)
    g = np.random.default_rng(11)
    seen = set()
    for prod in range(int(model["n_products"])):
        law = int(elaw[prod])
        if law == abi.OMCGPU_ELAW_NONE:
            continue
        tabs = []
        ok = True
        for k in range(2 if law == abi.OMCGPU_ELAW_WATT else (0 if law == abi.OMCGPU_ELAW_LEVEL else 1)):
            f = int(pfn[prod, k])
            if fn_type[f] != abi.OMCGPU_FN_TABULATED1D:
                ok = False
                break
            d = data[fn_ofs[f]:fn_ofs[f + 1]]
            tabs.append((np.ascontiguousarray(d[:d.size // 2]), np.ascontiguousarray(d[d.size // 2:])))
        if not ok:
            continue
        while len(tabs) < 2:
            tabs.append((np.zeros(1), np.zeros(1)))
        n = 4000
        lo = max(par[prod, 0] * 1.0001, 1e-3) + 1.0          # above the threshold / restriction energy
        E_in = np.ascontiguousarray(np.exp(g.uniform(np.log(lo), np.log(1.9e7), n)))
        seeds = g.integers(1, 2**63, n, dtype=np.uint64)
        E_r = np.zeros(n); mu_r = np.zeros(n); s_r = np.zeros(n, np.uint64)
        rc = L.ref_sample_energy_law(C.c_int(law), C.c_double(par[prod, 0]), C.c_double(par[prod, 1]),
                                     p(tabs[0][0]), p(tabs[0][1]), C.c_int(tabs[0][0].size),
                                     p(tabs[1][0]), p(tabs[1][1]), C.c_int(tabs[1][0].size),
                                     C.c_int64(n), p(E_in), p(seeds), p(E_r), p(mu_r), p(s_r))
        assert rc == 0, L.ref_geom_error()
        E_o, mu_o, s_o = o.sample_product(prod, E_in, seeds)
        assert np.array_equal(mu_o, mu_r) and np.array_equal(E_o, E_r) and np.array_equal(s_o, s_r), (prod, law)
        seen.add(law)
    assert abi.OMCGPU_ELAW_WATT in seen and len(seen) >= 2, seen


@needs_ref
def test_free_gas_target_velocity_equals_reference():
    """sample_cxs_target_velocity of the reference's src/physics.cpp (compiled unmodified):
    the rejection loop of the free-gas kernel consumes the same random numbers and returns the
    identical target velocity -- light and heavy targets, thermal to epithermal energies"""
    import ctypes as C
    model, box, of, libx = cases.build_ce(n_grid=500)
    L = C.CDLL(ref_geometry.LIB)
    o = Oracle(model); o.load()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    g = np.random.default_rng(17)
    n = 20000
    kT = 0.0253
    for awr in (0.9992, 15.86, 236.0):
        E = np.ascontiguousarray(np.exp(g.uniform(np.log(1e-4), np.log(400 * kT), n)))
        u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        u = np.ascontiguousarray(u)
        seeds = g.integers(1, 2**63, n, dtype=np.uint64)
        out = []
        for fn in (L.ref_sample_cxs_target_velocity, o.lib.orc_sample_cxs_target_velocity):
            v = np.zeros((n, 3)); s = np.zeros(n, np.uint64)
            assert fn(C.c_double(awr), C.c_double(kT), C.c_int64(n), p(E), p(u), p(seeds), p(v), p(s)) == 0
            out.append((v, s))
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]), awr


@needs_ref
def test_elastic_scattering_equals_reference():
    """elastic_scatter of the reference's src/physics.cpp (target velocity with its threshold
    for a target at rest, CM <-> LAB transforms, isotropic CM angle, rotate_angle) against the
    oracle's restatement for every nuclide of the synthetic library: identical outgoing
    energy, direction, lab cosine and seed -- thermal energies (free gas) up to 1 MeV"""
    import ctypes as C
    model, box, of, libx = cases.build_ce(n_grid=500)
    L = C.CDLL(ref_geometry.LIB)
    L.ref_geom_error.restype = C.c_char_p
    o = Oracle(model); o.load()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    g = np.random.default_rng(23)
    n = 20000
    kT = 0.0253
    awrs = np.asarray(model["nuc_awr"])
    for nuc in range(int(model["n_nuclides"])):
        E = np.ascontiguousarray(np.exp(g.uniform(np.log(1e-4), np.log(1e6), n)))
        u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        u = np.ascontiguousarray(u)
        seeds = g.integers(1, 2**63, n, dtype=np.uint64)
        res = []
        for fn, first in ((L.ref_elastic_scatter, C.c_double(float(awrs[nuc]))), (o.lib.orc_elastic_scatter, C.c_int(nuc))):
            Eo = np.zeros(n); uo = np.zeros((n, 3)); mu = np.zeros(n); so = np.zeros(n, np.uint64)
            rc = fn(first, C.c_double(kT), C.c_int64(n), p(E), p(u), p(seeds), p(Eo), p(uo), p(mu), p(so))
            assert rc == 0, L.ref_geom_error()
            res.append((Eo, uo, mu, so))
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), (nuc, float(awrs[nuc]))
    # (afterwards the reference library holds the one hand-made nuclide; ReferenceXS reloads its own)


def _reaction_into_reference(L, model, nuc, rx, fissionable):
    """hand the products of reaction `rx` (global index) of nuclide `nuc` to the reference
    library's hand-made Reaction; returns False when a law's parameter is not a table"""
    import ctypes as C
    from openmc_b200 import _abi as abi
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn_ofs = np.asarray(model["fn_ofs"]); fn_type = np.asarray(model["fn_type"]); data = np.asarray(model["fn_data"])
    elaw = np.asarray(model["prod_elaw"]); par = np.asarray(model["prod_par"]).reshape(-1, 2)
    pfn = np.asarray(model["prod_fn"]).reshape(-1, 2)
    fn = lambda f: np.ascontiguousarray(data[fn_ofs[f]:fn_ofs[f + 1]])
    p0, p1 = int(model["rx_prod_ofs"][rx]), int(model["rx_prod_ofs"][rx + 1])
    tn = int(model["nuc_total_nu_fn"][nuc]) if fissionable else -1
    tn_d = fn(tn) if tn >= 0 else np.zeros(1)
    rc = L.ref_rx_begin(C.c_int(int(model["rx_scatter_in_cm"][rx])), C.c_double(float(model["nuc_awr"][nuc])),
                        C.c_double(0.0253), C.c_int(int(model["nuc_n_precursor"][nuc])), C.c_int(int(fissionable)),
                        C.c_int(int(tn >= 0 and fn_type[tn] == abi.OMCGPU_FN_POLYNOMIAL)), p(tn_d),
                        C.c_int(tn_d.size if tn >= 0 else 0), C.c_int(p1 - p0), C.c_double(float(model["energy_max"])))
    assert rc == 0, L.ref_geom_error()
    return ref_geometry.products_into_reference(L, model, rx)


@needs_ref
def test_inelastic_scattering_and_fission_emission_equal_reference():
    """inelastic_scatter and sample_fission_neutron of the reference's src/physics.cpp (compiled
    unmodified) on every inelastic and fission reaction of the synthetic library: outgoing
    energy, direction, lab cosine, weight (yield), delayed group, emission time and the seed
    afterwards are identical to the oracle's restatement -- Nuclide::nu total / delayed, the
    precursor-group walk, ReactionProduct::sample, CM -> LAB, rotate_angle"""
    import ctypes as C
    model, box, of, libx = cases.build_ce(n_grid=500)
    L = C.CDLL(ref_geometry.LIB)
    L.ref_geom_error.restype = C.c_char_p
    o = Oracle(model); o.load()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    g = np.random.default_rng(29)
    n = 8000
    n_inel = n_fis = n_delayed = 0
    par = np.asarray(model["prod_par"]).reshape(-1, 2)
    for nuc in range(int(model["n_nuclides"])):
        r0, r1 = int(model["nuc_rx_ofs"][nuc]), int(model["nuc_rx_ofs"][nuc + 1])
        fis = r0 + int(model["nuc_fission_rx"][nuc]) if int(model["nuc_fissionable"][nuc]) else -1
        for rx in range(r0, r1):
            is_fis = rx == fis
            if not is_fis and not int(model["rx_inelastic"][rx]):
                continue
            if not _reaction_into_reference(L, model, nuc, rx, is_fis):
                continue
            prod0 = int(model["rx_prod_ofs"][rx])
            n = 60000 if is_fis else 8000          # enough histories for a few hundred delayed neutrons
            lo = max(par[prod0, 0] * 1.0001, 1e-3) + 1.0 if not is_fis else 1e-4
            E = np.ascontiguousarray(np.exp(g.uniform(np.log(lo), np.log(1.9e7), n)))
            u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
            u = np.ascontiguousarray(u)
            seeds = g.integers(1, 2**63, n, dtype=np.uint64)
            res = []
            if is_fis:
                for fn, head in ((L.ref_rx_fission_neutron, ()), (o.lib.orc_fission_neutron, (C.c_int(nuc), C.c_int(rx)))):
                    Eo = np.zeros(n); uo = np.zeros((n, 3)); grp = np.zeros(n, np.int32); t = np.zeros(n)
                    so = np.zeros(n, np.uint64)
                    rc = fn(*head, C.c_int64(n), p(E), p(u), p(seeds), p(Eo), p(uo), p(grp), p(t), p(so))
                    assert rc == 0, L.ref_geom_error()
                    res.append((Eo, uo, grp, t, so))
                n_fis += 1
                n_delayed += int((res[0][2] > 0).sum())
            else:
                for fn, head in ((L.ref_rx_inelastic, ()), (o.lib.orc_inelastic_scatter, (C.c_int(nuc), C.c_int(rx)))):
                    Eo = np.zeros(n); uo = np.zeros((n, 3)); mu = np.zeros(n); w = np.zeros(n); so = np.zeros(n, np.uint64)
                    rc = fn(*head, C.c_int64(n), p(E), p(u), p(seeds), p(Eo), p(uo), p(mu), p(w), p(so))
                    assert rc == 0, L.ref_geom_error()
                    res.append((Eo, uo, mu, w, so))
                n_inel += 1
            for a, b in zip(res[0], res[1]):
                assert np.array_equal(a, b), (nuc, rx, is_fis)
    assert n_inel > 0 and n_fis > 0 and n_delayed > 50, (n_inel, n_fis, n_delayed)
    print(f"reactions compared: {n_inel} inelastic, {n_fis} fission ({n_delayed} delayed neutrons)")


@needs_ref
@pytest.mark.parametrize("survival", [0, 1])
def test_whole_collision_equals_reference(survival):
    """collision() of the reference's src/physics.cpp (compiled unmodified, with reaction.cpp and
    reaction_product.cpp) on hand-made nuclides and materials that carry the synthetic library:
    sample_nuclide, sample_fission + create_fission_sites, absorption, the elastic / inelastic
    channel walk and the sampled reaction leave every neutron with the oracle's energy,
    direction, weight, k-absorption score and seed, and bank the identical fission sites
    (count, energy, direction, delayed group, emission time) -- every material, 1e-4 eV to 19 MeV;
    analog absorption, and survival biasing with russian_roulette (src/physics_common.cpp) on
    neutrons that start just above the weight cutoff"""
    model, box, of, libx = cases.build_ce(n_grid=500)
    wgt0 = 0.3 if survival else 1.0
    ref = ref_geometry.ReferenceXS(model)
    o = Oracle(model); o.load()
    mat, E, u, seeds = ref_geometry.collision_probes(ref, 60000, seed=31)
    n = E.size
    model.struct.survival_biasing = survival      # the oracle reads the live struct; the model is cached
    try:
        ref.load_collision_physics(survival)
        r = ref.collisions(mat, E, u, seeds, 1.1, wgt0)
        q = ref_geometry.oracle_collisions(o, mat, E, u, seeds, 1.1, wgt0)
    finally:
        model.struct.survival_biasing = 0         # ... by cases.build_ce: put it back
        ref.unload_collision_physics()
    for name in ref_geometry.COLLISION_FIELDS:
        assert np.array_equal(r[name], q[name]), name
    # every channel was exercised
    mt, inuc = r["mt"], r["nuclide"]
    absorbed = int((r["wgt"] == 0).sum()); n_sites = int(r["site_ofs"][-1])
    elastic = int((mt == 2).sum()); inelastic = int(((mt != 2) & (mt != 18) & (r["wgt"] > 0)).sum())
    print(f"collisions: {n}, absorbed {absorbed}, fission sites {n_sites}, elastic {elastic}, inelastic {inelastic}, "
          f"nuclides hit {np.unique(inuc).size}")
    assert absorbed > 100 and n_sites > 100 and elastic > 1000 and inelastic > 100
    if survival:
        rouletted_up = int((r["wgt"] == float(model["weight_survive"])).sum())
        assert rouletted_up > 100, rouletted_up
    assert np.unique(inuc).size == int(model["n_nuclides"])
