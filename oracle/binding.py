"""ctypes binding of liboracle.so (the CPU restatement) and, when built, of
oracle/_ref/libref_helpers.so (the reference's own helper TUs compiled verbatim).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from openmc_b200 import _abi as abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_helpers.so")
REF_GEOM_LIB = os.path.join(HERE, "_ref", "libref_geom.so")


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(HERE, "oracle.cpp")
    hdr = abi.HEADER
    stale = (not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src)
             or os.path.getmtime(LIB) < os.path.getmtime(hdr))
    have_ref = os.path.isdir("/root/reference/src")
    shim = os.path.join(HERE, "ref_geom_shim.cpp")
    full = os.path.join(HERE, "_ref", "openmc")
    h5shim = os.path.join(HERE, "h5shim", "h5shim.cpp")
    ref_stale = have_ref and (not os.path.exists(REF_LIB) or not os.path.exists(REF_GEOM_LIB)
                              or not os.path.exists(full) or os.path.getmtime(full) < os.path.getmtime(h5shim)
                              or os.path.getmtime(REF_GEOM_LIB) < os.path.getmtime(shim)
                              or os.path.getmtime(REF_GEOM_LIB) < os.path.getmtime(os.path.join(HERE, "Makefile")))
    if force or stale or ref_stale:
        subprocess.check_call(["make", "-s", "-C", HERE])


class Settings(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("n_batches", C.c_int32), ("n_inactive", C.c_int32),
                ("gen_per_batch", C.c_int32), ("reserved", C.c_int32), ("master_seed", C.c_int64),
                ("stride", C.c_uint64)]


class Results(C.Structure):
    _fields_ = [("k_combined", C.c_double * 2), ("keff", C.c_double), ("keff_std", C.c_double),
                ("global_sum", C.c_double * 4), ("global_sum_sq", C.c_double * 4),
                ("n_realizations", C.c_int64), ("n_lost", C.c_int64), ("n_events", C.c_int64),
                ("n_xs_lookups", C.c_int64), ("seconds_active", C.c_double)]


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class Oracle:
    def __init__(self, model: abi.ModelHandle):
        build()
        self.lib = C.CDLL(LIB)
        self.model = model
        self.lib.orc_load(model.byref())
        self.lib.orc_evaluate_legendre.restype = C.c_double
        self.lib.orc_maxwell_spectrum.restype = C.c_double
        self.lib.orc_watt_spectrum.restype = C.c_double
        self.lib.orc_lower_bound_index.restype = C.c_int64
        self.lib.orc_comb_indices.restype = C.c_int64
        self.lib.orc_last_error.restype = C.c_char_p

    def load(self):
        self.lib.orc_load(self.model.byref())

    # ---- probes --------------------------------------------------------------
    def future_seed(self, n, seed):
        n = np.ascontiguousarray(n, np.uint64); seed = np.ascontiguousarray(seed, np.uint64)
        out = np.zeros_like(n)
        self.lib.orc_future_seed(_p(n, C.c_uint64), _p(seed, C.c_uint64), C.c_int64(n.size),
                                 _p(out, C.c_uint64))
        return out

    def prn(self, seed, count):
        seed = np.ascontiguousarray(seed, np.uint64)
        out = np.zeros((seed.size, count))
        self.lib.orc_prn(_p(seed, C.c_uint64), C.c_int64(seed.size), C.c_int32(count),
                         _p(out, C.c_double))
        return out

    def find_cell(self, xyz, uvw=None):
        self.load()
        xyz = np.ascontiguousarray(xyz, np.float64); n = xyz.shape[0]
        cell = np.zeros(n, np.int32); inst = np.zeros(n, np.int32)
        mat = np.zeros(n, np.int32); nco = np.zeros(n, np.int32)
        up = None
        if uvw is not None:
            uvw = np.ascontiguousarray(uvw, np.float64); up = _p(uvw, C.c_double)
        self.lib.orc_find_cell(_p(xyz, C.c_double), up, C.c_int64(n), _p(cell, C.c_int32),
                               _p(inst, C.c_int32), _p(mat, C.c_int32), _p(nco, C.c_int32))
        return cell, inst, mat, nco

    def distance_to_boundary(self, xyz, uvw):
        self.load()
        xyz = np.ascontiguousarray(xyz, np.float64); uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        d = np.zeros(n); s = np.zeros(n, np.int32); lv = np.zeros(n, np.int32)
        lt = np.zeros((n, 3), np.int32)
        self.lib.orc_distance_to_boundary(_p(xyz, C.c_double), _p(uvw, C.c_double), C.c_int64(n),
                                          _p(d, C.c_double), _p(s, C.c_int32), _p(lv, C.c_int32),
                                          _p(lt, C.c_int32))
        return d, s, lv, lt

    def calculate_xs(self, E, material, max_nuc=0):
        self.load()
        E = np.ascontiguousarray(E, np.float64); material = np.ascontiguousarray(material, np.int32)
        n = E.size
        out = np.zeros((n, 4)); ig = np.zeros((n, max(max_nuc, 1)), np.int32)
        self.lib.orc_calculate_xs(_p(E, C.c_double), _p(material, C.c_int32), C.c_int64(n),
                                  _p(out, C.c_double), _p(ig, C.c_int32) if max_nuc else None,
                                  C.c_int32(max_nuc))
        return out, ig

    def geom_track(self, xyz, uvw, steps):
        self.load()
        xyz = np.ascontiguousarray(xyz, np.float64); uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        cid = np.zeros((n, steps), np.int32); mat = np.zeros((n, steps), np.int32)
        nco = np.zeros((n, steps), np.int32); d = np.zeros((n, steps))
        self.lib.orc_geom_track(C.c_int64(n), _p(xyz, C.c_double), _p(uvw, C.c_double), C.c_int(steps),
                                _p(cid, C.c_int32), _p(mat, C.c_int32), _p(nco, C.c_int32), _p(d, C.c_double))
        return cid, mat, nco, d

    def sample_product(self, prod, E_in, seeds):
        self.load()
        E_in = np.ascontiguousarray(E_in, np.float64); seeds = np.ascontiguousarray(seeds, np.uint64)
        E_out = np.zeros_like(E_in); mu = np.zeros_like(E_in); s_out = np.zeros_like(seeds)
        self.lib.orc_sample_product(C.c_int(prod), C.c_int64(E_in.size), _p(E_in, C.c_double), _p(seeds, C.c_uint64),
                                    _p(E_out, C.c_double), _p(mu, C.c_double), _p(s_out, C.c_uint64))
        return E_out, mu, s_out

    def eval_function(self, fn, x):
        self.load()
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        self.lib.orc_eval_function(C.c_int(fn), C.c_int64(x.size), _p(x, C.c_double), _p(y, C.c_double))
        return y

    def mesh_bins_crossed(self, mesh, r0, r1, u, cap_per_track=512):
        """(ofs, bins, length fractions) of StructuredMesh::bins_crossed for tracks r0 -> r1"""
        self.load()
        r0 = np.ascontiguousarray(r0, np.float64); r1 = np.ascontiguousarray(r1, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        n = r0.shape[0]
        cap = n * cap_per_track
        ofs = np.zeros(n + 1, np.int64); bins = np.zeros(cap, np.int32); lengths = np.zeros(cap)
        self.lib.orc_mesh_bins_crossed.restype = C.c_int64
        tot = self.lib.orc_mesh_bins_crossed(C.c_int(mesh), C.c_int64(n), _p(r0, C.c_double), _p(r1, C.c_double),
                                             _p(u, C.c_double), _p(ofs, C.c_int64), _p(bins, C.c_int32),
                                             _p(lengths, C.c_double), C.c_int64(cap))
        assert tot >= 0
        return ofs, bins[:tot], lengths[:tot]

    # ---- transport -----------------------------------------------------------
    def sample_initial_source(self, n, box, only_fissionable=False, master_seed=1,
                              stride=152917, id_offset=0, watt=(0.988e6, 2.249e-6)):
        self.load()
        bank = np.zeros(n, abi.site_dtype)
        box = np.ascontiguousarray(box, np.float64)
        rc = self.lib.orc_sample_initial_source(
            C.c_int64(n), C.c_int64(id_offset), C.c_int64(0), C.c_int64(master_seed),
            C.c_uint64(stride), _p(box, C.c_double), C.c_int(int(only_fissionable)),
            C.c_double(watt[0]), C.c_double(watt[1]), bank.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return bank

    def transport_generation(self, gen_in: abi.GenIn, source_bank, fission_capacity=None):
        self.load()
        n = int(gen_in.n_particles)
        cap = 3 * n if fission_capacity is None else fission_capacity
        fb = np.zeros(cap, abi.site_dtype)
        out = abi.GenOut()
        src = np.ascontiguousarray(source_bank)
        self.lib.orc_transport_generation(C.byref(gen_in), src.ctypes.data_as(C.c_void_p),
                                          fb.ctypes.data_as(C.c_void_p), C.c_int64(cap),
                                          C.byref(out))
        return fb[:out.n_fission], out

    def tally_results(self, t, n):
        v = np.zeros(n)
        self.lib.orc_tally_results(C.c_int32(t), _p(v, C.c_double), C.c_int64(n))
        return v

    def comb_indices(self, start, n_local, total, n_particles, seed_id, master_seed=1,
                     stride=152917):
        idx = np.zeros(n_particles + 2, np.int64)
        n = self.lib.orc_comb_indices(C.c_int64(start), C.c_int64(n_local), C.c_int64(total),
                                      C.c_int64(n_particles), C.c_int64(seed_id),
                                      C.c_int64(master_seed), C.c_uint64(stride),
                                      _p(idx, C.c_int64), C.c_int64(idx.size))
        return idx[:n]

    def run_eigenvalue(self, source_bank, n_batches, n_inactive, gen_per_batch=1, master_seed=1,
                       stride=152917):
        self.load()
        n = source_bank.size
        st = Settings(n, n_batches, n_inactive, gen_per_batch, 0, master_seed, stride)
        res = Results()
        kgen = np.zeros(n_batches * gen_per_batch)
        nt = int(self.model["tally_result_ofs"][self.model["n_tallies"]]) if self.model["n_tallies"] else 0
        tsum = np.zeros(max(nt, 1)); tsq = np.zeros(max(nt, 1))
        bank = np.ascontiguousarray(source_bank.copy())
        rc = self.lib.orc_run_eigenvalue(C.byref(st), bank.ctypes.data_as(C.c_void_p), C.byref(res),
                                         _p(kgen, C.c_double), _p(tsum, C.c_double),
                                         _p(tsq, C.c_double))
        if rc:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return dict(k_combined=(res.k_combined[0], res.k_combined[1]), keff=res.keff,
                    keff_std=res.keff_std, k_generation=kgen, tally_sum=tsum[:nt],
                    tally_sum_sq=tsq[:nt], global_sum=np.array(res.global_sum[:]),
                    global_sum_sq=np.array(res.global_sum_sq[:]),
                    n_realizations=res.n_realizations, n_lost=res.n_lost, n_events=res.n_events,
                    n_xs_lookups=res.n_xs_lookups, source_bank=bank)


class Reference:
    """The reference's own helper functions (oracle/_ref, built from /root/reference)."""

    def __init__(self):
        build()
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(REF_LIB)
        L = self.lib = C.CDLL(REF_LIB)
        L.ref_prn.restype = C.c_double
        L.ref_future_seed.restype = C.c_uint64
        L.ref_future_seed.argtypes = [C.c_uint64, C.c_uint64]
        L.ref_init_seed.restype = C.c_uint64
        L.ref_init_seed.argtypes = [C.c_int64, C.c_int]
        L.ref_evaluate_legendre.restype = C.c_double
        L.ref_maxwell_spectrum.restype = C.c_double
        L.ref_watt_spectrum.restype = C.c_double
        L.ref_lower_bound_index.restype = C.c_int64
