"""Doorway to the reference's OWN geometry code for the tests: emits a geometry.xml
text (format: docs/source/io_formats/geometry.rst, read by src/surface.cpp:1195+,
src/cell.cpp:1170+, src/lattice.cpp:162-232) from this repo's small model description
and asks oracle/_ref/libref_geom.so -- the unmodified reference translation units --
where points are and how far their next boundary is."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_geom.so")


from oracle.refrun import LATTICE_ID_BASE, _fmt, geometry_xml  # noqa: E402,F401  (the emitters live with the other XML writers)


class ReferenceGeometry:
    def __init__(self, m):
        if not os.path.exists(LIB):
            raise FileNotFoundError(LIB)
        self.lib = L = C.CDLL(LIB)
        L.ref_geom_error.restype = C.c_char_p
        ids = np.array([mat.id for mat in m.materials], np.int32)
        rc = L.ref_geom_load(geometry_xml(m).encode(), ids.ctypes.data_as(C.c_void_p), C.c_int(ids.size))
        if rc:
            raise RuntimeError(L.ref_geom_error().decode())
        c = np.zeros(6, np.int32)
        L.ref_geom_counts(c.ctypes.data_as(C.c_void_p))
        self.n_surfaces, self.n_cells, self.n_universes, self.n_lattices, self.n_coord_levels, self.root = \
            (int(x) for x in c)

    def find_cell(self, xyz, uvw):
        xyz = np.ascontiguousarray(xyz, np.float64)
        uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        cid = np.zeros(n, np.int32); inst = np.zeros(n, np.int32)
        mat = np.zeros(n, np.int32); nco = np.zeros(n, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_geom_find_cell(C.c_int64(n), p(xyz), p(uvw), p(cid), p(inst), p(mat), p(nco)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return cid, inst, mat, nco

    def track(self, xyz, uvw, steps):
        """straight-line tracking with the reference's crossing logic (ref_geom_track)"""
        xyz = np.ascontiguousarray(xyz, np.float64); uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        cid = np.zeros((n, steps), np.int32); mat = np.zeros((n, steps), np.int32)
        nco = np.zeros((n, steps), np.int32); d = np.zeros((n, steps))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_geom_track(C.c_int64(n), p(xyz), p(uvw), C.c_int(steps), p(cid), p(mat), p(nco), p(d)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return cid, mat, nco, d

    def reflect(self, token, xyz, uvw):
        """new directions from the reference's ReflectiveBC for particles on surface `token`"""
        token = np.ascontiguousarray(token, np.int32)
        xyz = np.ascontiguousarray(xyz, np.float64); uvw = np.ascontiguousarray(uvw, np.float64)
        new = np.zeros_like(uvw)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_geom_reflect(C.c_int64(token.size), p(token), p(xyz), p(uvw), p(new)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return new

    def periodic(self, token, xyz, uvw):
        """what the reference's periodic BC does to particles that struck surface `token`
        (signed index + 1): new positions and new surface tokens (0: not a periodic surface)"""
        token = np.ascontiguousarray(token, np.int32)
        xyz = np.ascontiguousarray(xyz, np.float64)
        uvw = np.ascontiguousarray(uvw, np.float64)
        new_xyz = np.zeros_like(xyz); new_token = np.zeros_like(token)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_geom_periodic(C.c_int64(token.size), p(token), p(xyz), p(uvw), p(new_xyz), p(new_token)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return new_xyz, new_token

    def distance_to_boundary(self, xyz, uvw):
        xyz = np.ascontiguousarray(xyz, np.float64)
        uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        d = np.zeros(n); s = np.zeros(n, np.int32); lv = np.zeros(n, np.int32)
        lt = np.zeros((n, 3), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_geom_distance(C.c_int64(n), p(xyz), p(uvw), p(d), p(s), p(lv), p(lt)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return d, s, lv, lt


class ReferenceXS:
    """The reference's Nuclide::init_grid + Nuclide::calculate_xs (src/nuclide.cpp:502-533,
    616-839, compiled unmodified) on the pointwise tables of a flattened CE model."""

    def __init__(self, model):
        if not os.path.exists(LIB):
            raise FileNotFoundError(LIB)
        self.lib = L = C.CDLL(LIB)
        L.ref_geom_error.restype = C.c_char_p
        self.model = model
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.nn = int(model["n_nuclides"])
        self.ofs = np.ascontiguousarray(model["nuc_grid_ofs"], np.int64)
        en = np.ascontiguousarray(model["nuc_energy"], np.float64)
        xs = np.ascontiguousarray(model["nuc_xs"], np.float64)
        fis = np.ascontiguousarray(model["nuc_fissionable"], np.int32)
        kT = np.ascontiguousarray(model["nuc_kT"], np.float64)
        self.n_log_bins = int(model["n_log_bins"])
        self.e_min, self.e_max = float(model["energy_min"]), float(model["energy_max"])
        if L.ref_xs_load(C.c_int(self.nn), p(self.ofs), p(en), p(xs), p(fis), p(kT), C.c_int(self.n_log_bins),
                         C.c_double(self.e_min), C.c_double(self.e_max)):
            raise RuntimeError(L.ref_geom_error().decode())

    def in_range(self, E):
        """energies the reference can be asked about: outside [E_min, E_max) its hash
        lookup indexes past the table (src/nuclide.cpp:732-733; transport never gets there)"""
        E = np.asarray(E, float)
        return np.ascontiguousarray(E[(E >= self.e_min) & (E < self.e_max * (1 - 1e-9))])

    def grid_index(self, nuc):
        gi = np.zeros(self.n_log_bins + 1, np.int32)
        assert self.lib.ref_xs_grid_index(C.c_int(nuc), gi.ctypes.data_as(C.c_void_p), C.c_int64(gi.size)) == 0
        return gi

    def micro(self, nuc, E, sqrtkT=0.15906):
        E = np.ascontiguousarray(E, np.float64)
        out = np.zeros((E.size, 4)); ig = np.zeros(E.size, np.int32); fr = np.zeros(E.size)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.lib.ref_xs_micro(C.c_int(int(nuc)), C.c_int64(E.size), p(E), C.c_double(sqrtkT), p(out), p(ig), p(fr)):
            raise RuntimeError(self.lib.ref_geom_error().decode())
        return out, ig, fr

    def macro(self, mat, E):
        """sum over the material's nuclides in order (src/material.cpp:893-899); returns the
        sums [n][4] and the grid indices [n][n_nuclides of the material]"""
        m = self.model
        k0, k1 = int(m["mat_nuc_ofs"][mat]), int(m["mat_nuc_ofs"][mat + 1])
        tot = np.zeros((np.size(E), 4)); igs = []
        for k in range(k0, k1):
            out, ig, _ = self.micro(int(m["mat_nuclide"][k]), E)
            tot += float(m["mat_atom_density"][k]) * out
            igs.append(ig)
        return tot, (np.stack(igs, 1) if igs else np.zeros((np.size(E), 0), np.int32))


class ReferenceMesh:
    """The reference's RegularMesh (src/mesh.cpp, compiled unmodified) built from <mesh> elements."""

    def __init__(self, meshes):
        if not os.path.exists(LIB):
            raise FileNotFoundError(LIB)
        self.lib = L = C.CDLL(LIB)
        L.ref_geom_error.restype = C.c_char_p
        L.ref_mesh_bins_crossed.restype = C.c_int64
        x = ["<tallies>"]
        for m in meshes:
            x.append(f'  <mesh id="{m.id}"><dimension>{" ".join(str(int(d)) for d in m.dimension)}</dimension>'
                     f"<lower_left>{_fmt(m.lower_left)}</lower_left><upper_right>{_fmt(m.upper_right)}</upper_right></mesh>")
        x.append("</tallies>")
        if L.ref_mesh_load("\n".join(x).encode()):
            raise RuntimeError(L.ref_geom_error().decode())

    def bins_crossed(self, mesh, r0, r1, u, cap_per_track=512):
        r0 = np.ascontiguousarray(r0, np.float64); r1 = np.ascontiguousarray(r1, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        n = r0.shape[0]
        cap = n * cap_per_track
        ofs = np.zeros(n + 1, np.int64); bins = np.zeros(cap, np.int32); lengths = np.zeros(cap)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        tot = self.lib.ref_mesh_bins_crossed(C.c_int(mesh), C.c_int64(n), p(r0), p(r1), p(u), p(ofs), p(bins),
                                             p(lengths), C.c_int64(cap))
        if tot < 0:
            raise RuntimeError(self.lib.ref_geom_error().decode() or "capacity")
        return ofs, bins[:tot], lengths[:tot]


def products_into_reference(L, model, rx, allow_no_law=False):
    """fills the products of the library's current hand-made Reaction from reaction `rx`
    (global index) of the model; False when a law's parameter is not a table"""
    from openmc_b200 import _abi as abi
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn_ofs = np.asarray(model["fn_ofs"]); fn_type = np.asarray(model["fn_type"]); data = np.asarray(model["fn_data"])
    elaw = np.asarray(model["prod_elaw"]); par = np.asarray(model["prod_par"]).reshape(-1, 2)
    pfn = np.asarray(model["prod_fn"]).reshape(-1, 2)
    fn = lambda f: np.ascontiguousarray(data[fn_ofs[f]:fn_ofs[f + 1]])
    p0, p1 = int(model["rx_prod_ofs"][rx]), int(model["rx_prod_ofs"][rx + 1])
    for i, prod in enumerate(range(p0, p1)):
        law = int(elaw[prod])
        if law == abi.OMCGPU_ELAW_NONE and not allow_no_law:
            return False
        tabs = []
        for k in range(2 if law == abi.OMCGPU_ELAW_WATT else (0 if law in (abi.OMCGPU_ELAW_LEVEL, abi.OMCGPU_ELAW_NONE) else 1)):
            f = int(pfn[prod, k])
            if fn_type[f] != abi.OMCGPU_FN_TABULATED1D:
                return False
            d = fn(f)
            tabs.append((np.ascontiguousarray(d[:d.size // 2]), np.ascontiguousarray(d[d.size // 2:])))
        while len(tabs) < 2:
            tabs.append((np.zeros(1), np.zeros(1)))
        y = int(model["prod_yield_fn"][prod]); y_d = fn(y)
        rc = L.ref_rx_product(C.c_int(i), C.c_int(law), C.c_double(par[prod, 0]), C.c_double(par[prod, 1]),
                              p(tabs[0][0]), p(tabs[0][1]), C.c_int(tabs[0][0].size),
                              p(tabs[1][0]), p(tabs[1][1]), C.c_int(tabs[1][0].size),
                              C.c_int(int(fn_type[y] == abi.OMCGPU_FN_POLYNOMIAL)), p(y_d), C.c_int(y_d.size),
                              C.c_int(int(model["prod_emission"][prod])), C.c_double(float(model["prod_decay_rate"][prod])))
        assert rc == 0, L.ref_geom_error()
    return True


COLLISION_FIELDS = "E u wgt k_abs seed site_ofs site_E site_u site_group site_time".split()


def collision_probes(ref, n, seed):
    """(material, energy, direction, seed) of n neutrons about to collide"""
    g = np.random.default_rng(seed)
    E = ref.in_range(np.exp(g.uniform(np.log(1e-4), np.log(1.9e7), n)))
    n = E.size
    mat = np.ascontiguousarray(g.integers(0, int(ref.model["n_materials"]), n), np.int32)
    u = g.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    return mat, E, np.ascontiguousarray(u), g.integers(1, 2**63, n, dtype=np.uint64)


def _collision_buffers(n):
    cap = 8 * n
    return dict(E=np.zeros(n), u=np.zeros((n, 3)), wgt=np.zeros(n), k_abs=np.zeros(n), seed=np.zeros(n, np.uint64),
                site_ofs=np.zeros(n + 1, np.int64), site_E=np.zeros(cap), site_u=np.zeros((cap, 3)),
                site_group=np.zeros(cap, np.int32), site_time=np.zeros(cap)), cap


def _trim_sites(b, tot):
    for k in ("site_E", "site_u", "site_group", "site_time"):
        b[k] = b[k][:tot]
    return b


def oracle_collisions(o, mat, E, u, seeds, k_eff, wgt0):
    """the oracle's collision_ce on the same neutrons (oracle/oracle.cpp: orc_collision)"""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = E.size
    b, cap = _collision_buffers(n)
    o.lib.orc_collision.restype = C.c_int64
    tot = o.lib.orc_collision(C.c_int64(n), p(mat), p(E), p(u), p(seeds), C.c_double(k_eff), C.c_double(wgt0),
                              p(b["E"]), p(b["u"]), p(b["wgt"]), p(b["k_abs"]), p(b["seed"]), p(b["site_ofs"]),
                              p(b["site_E"]), p(b["site_u"]), p(b["site_group"]), p(b["site_time"]), C.c_int64(cap))
    assert tot >= 0
    return _trim_sites(b, tot)


def _load_collision_physics(self, survival=0):
    """hand-made Reaction / ReactionProduct / Material objects of the reference library that
    carry the model's reactions, products and compositions (after ref_xs_load)"""
    from openmc_b200 import _abi as abi
    model, L = self.model, self.lib
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn_ofs = np.asarray(model["fn_ofs"]); fn_type = np.asarray(model["fn_type"]); data = np.asarray(model["fn_data"])
    rx_xs_ofs = np.asarray(model["rx_xs_ofs"]); rx_xs = np.asarray(model["rx_xs"])
    for nuc in range(int(model["n_nuclides"])):
        fissionable = int(model["nuc_fissionable"][nuc])
        tn = int(model["nuc_total_nu_fn"][nuc]) if fissionable else -1
        tn_d = np.ascontiguousarray(data[fn_ofs[tn]:fn_ofs[tn + 1]]) if tn >= 0 else np.zeros(1)
        assert L.ref_nuc_physics(C.c_int(nuc), C.c_double(float(model["nuc_awr"][nuc])),
                                 C.c_int(int(model["nuc_n_precursor"][nuc])),
                                 C.c_int(int(tn >= 0 and fn_type[tn] == abi.OMCGPU_FN_POLYNOMIAL)), p(tn_d),
                                 C.c_int(tn_d.size if tn >= 0 else 0)) == 0
        r0, r1 = int(model["nuc_rx_ofs"][nuc]), int(model["nuc_rx_ofs"][nuc + 1])
        fis = r0 + int(model["nuc_fission_rx"][nuc]) if fissionable else -1
        for rx in range(r0, r1):
            v = np.ascontiguousarray(rx_xs[rx_xs_ofs[rx]:rx_xs_ofs[rx + 1]])
            n_prod = int(model["rx_prod_ofs"][rx + 1]) - int(model["rx_prod_ofs"][rx])
            rc = L.ref_nuc_reaction(C.c_int(nuc), C.c_int(int(model["rx_mt"][rx])), C.c_int(int(model["rx_scatter_in_cm"][rx])),
                                    C.c_int(int(model["rx_inelastic"][rx])), C.c_int(int(rx == fis)),
                                    C.c_int(int(model["rx_threshold"][rx])), p(v), C.c_int64(v.size), C.c_int(n_prod))
            assert rc == 0, L.ref_geom_error()
            assert products_into_reference(L, model, rx, allow_no_law=True), (nuc, rx)
    self._m = (np.ascontiguousarray(model["mat_nuc_ofs"], np.int32), np.ascontiguousarray(model["mat_nuclide"], np.int32),
               np.ascontiguousarray(model["mat_atom_density"], np.float64))
    assert L.ref_materials_load(C.c_int(int(model["n_materials"])), *[p(a) for a in self._m]) == 0, L.ref_geom_error()
    L.ref_collision_settings(C.c_int(int(survival)), C.c_double(float(model["weight_cutoff"])),
                             C.c_double(float(model["weight_survive"])))


def _unload_collision_physics(self):
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if getattr(self, "_m", None) is not None:
        self.lib.ref_materials_load(C.c_int(0), *[p(a) for a in self._m])     # leave the library as found
    self.lib.ref_collision_settings(C.c_int(0), C.c_double(0.25), C.c_double(1.0))


def _collisions(self, mat, E, u, seeds, k_eff, wgt0):
    """the reference's collision() (src/physics.cpp:43-168) on n neutrons; adds the sampled
    nuclide and MT to the fields of COLLISION_FIELDS"""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = E.size
    b, cap = _collision_buffers(n)
    b["nuclide"] = np.zeros(n, np.int32); b["mt"] = np.zeros(n, np.int32)
    self.lib.ref_collision.restype = C.c_int64
    tot = self.lib.ref_collision(C.c_int64(n), p(mat), p(E), p(u), p(seeds), C.c_double(k_eff), C.c_double(wgt0),
                                 p(b["E"]), p(b["u"]), p(b["wgt"]), p(b["k_abs"]), p(b["nuclide"]), p(b["mt"]),
                                 p(b["seed"]), p(b["site_ofs"]), p(b["site_E"]), p(b["site_u"]), p(b["site_group"]),
                                 p(b["site_time"]), C.c_int64(cap))
    if tot < 0:
        raise RuntimeError(self.lib.ref_geom_error().decode())
    return _trim_sites(b, tot)


ReferenceXS.load_collision_physics = _load_collision_physics
ReferenceXS.unload_collision_physics = _unload_collision_physics
ReferenceXS.collisions = _collisions
