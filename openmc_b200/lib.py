"""Host side above the C ABI.

``Engine`` is the ctypes binding of ``libomcgpu.so`` (include/omcgpu.h) -- the
same binding a maintainer would add to ``openmc/lib/core.py``.

``Simulation`` stands in for the *untouched* L4 driver of libopenmc
(``openmc_simulation_init`` / ``openmc_next_batch`` / ``finalize_generation`` /
``finalize_batch``: src/simulation.cpp:75-280, 359-571; k-eff statistics:
src/eigenvalue.cpp:50-84, 367-577; tally accumulation:
src/tallies/tally.cpp:821-878, 1059-1104) so that the device path can be driven
and checked end to end here, where the reference itself cannot be built.  Its
method names follow ``openmc.lib`` (openmc/lib/core.py: simulation_init,
next_batch, run, keff, global_tallies, source_bank, iter_batches).

The product path needs the CUDA library: there is no CPU fallback and loading
fails loudly when ``libomcgpu.so`` is missing.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("OMCGPU_LIB") or os.path.join(_HERE, "csrc", "libomcgpu.so")
# the same sources with the rarely needed features compiled in (csrc/build.py: hexagonal lattices,
# secondary neutrons): Engine.init loads it for the models omcgpu_needs_extended_build names
EXT_LIB = os.path.join(_HERE, "csrc", "libomcgpu_ext.so")

DEFAULT_SEED = 1            # include/openmc/random_lcg.h:17
DEFAULT_STRIDE = 152917     # include/openmc/random_lcg.h:18


class OmcGpuError(RuntimeError):
    pass


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class Engine:
    """ctypes binding of the C ABI in include/omcgpu.h."""

    def __init__(self, libpath: str | None = None):
        path = libpath or DEFAULT_LIB
        if not os.path.exists(path):
            raise OmcGpuError(
                f"{path} not found: build it with `python -m openmc_b200.csrc.build` "
                "(the device path has no CPU fallback)")
        self._default_path = path
        self._load(path)
        self.model = None

    def _load(self, path):
        self.path = path
        self.lib = L = C.CDLL(path)
        L.omcgpu_last_error.restype = C.c_char_p
        L.omcgpu_kernel_launches.restype = C.c_int64
        L.omcgpu_lost_particles.restype = C.c_int64

    # -- helpers -----------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise OmcGpuError(f"[{rc}] " + self.lib.omcgpu_last_error().decode())

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.omcgpu_kernel_launches())

    # -- lifecycle ---------------------------------------------------------------
    def init(self, model: abi.ModelHandle, device: int = 0, max_particles: int = 100000):
        self.model = model
        self.max_particles = int(max_particles)
        # the build that carries the rarely needed features, for the models that need them
        want = self._default_path
        if os.path.basename(self._default_path) == "libomcgpu.so":
            if self.path != self._default_path:
                self._load(self._default_path)
            if self.lib.omcgpu_needs_extended_build(model.byref()):
                want = os.path.join(os.path.dirname(self._default_path), "libomcgpu_ext.so")
                if not os.path.exists(want):
                    raise OmcGpuError(f"{want} not found: build it with `python -m openmc_b200.csrc.build`")
        if want != self.path:
            self._load(want)
        self._check(self.lib.omcgpu_init(model.byref(), C.c_int(device), C.c_int64(max_particles)))

    def finalize(self):
        self._check(self.lib.omcgpu_finalize())

    def lost_particles(self, capacity: int = 64):
        """{id, r, u, E} of the first lost particles of the last generation (what
        Particle::mark_as_lost reports, src/particle.cpp:762-785) -> array [n][8]"""
        rec = np.zeros((max(int(capacity), 1), 8))
        n = int(self.lib.omcgpu_lost_particles(rec.ctypes.data_as(C.c_void_p), C.c_int64(capacity)))
        if n < 0:
            self._check(-n)
        return rec[:n]

    def count_fission_sites(self, shape, lower_left, width):
        """Mesh::count_sites over the device fission bank of the last generation (Shannon entropy):
        -> (weights per mesh bin, sites_outside)"""
        shape = np.ascontiguousarray(shape, np.int32)
        ll = np.ascontiguousarray(lower_left, np.float64)
        w = np.ascontiguousarray(width, np.float64)
        counts = np.zeros(int(np.prod(shape)))
        outside = C.c_int32(0)
        self._check(self.lib.omcgpu_count_fission_sites(
            C.c_int32(shape.size), shape.ctypes.data_as(C.c_void_p), ll.ctypes.data_as(C.c_void_p),
            w.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), C.c_int64(counts.size),
            C.byref(outside)))
        return counts, bool(outside.value)

    def set_option(self, name: str, value: int):
        self._check(self.lib.omcgpu_set_option(name.encode(), C.c_int64(value)))

    @property
    def sweep_kind(self) -> int:
        """2: paired sweep, lookup + flight inside the crossing / collision kernels; 1: merged
        sweep with one lookup + flight kernel; 0: merged sweep with the stand-alone lookup kernel
        over the sorted list (many-nuclide models)"""
        return int(self.lib.omcgpu_sweep_kind())

    @property
    def fused_lookup(self) -> bool:
        return self.sweep_kind == 2

    # -- banks -------------------------------------------------------------------
    def set_source_bank(self, bank: np.ndarray):
        bank = np.ascontiguousarray(bank, dtype=abi.site_dtype)
        self._check(self.lib.omcgpu_set_source_bank(bank.ctypes.data_as(C.c_void_p),
                                                    C.c_int64(bank.size)))

    def get_source_bank(self, n: int, out: np.ndarray | None = None) -> np.ndarray:
        """device source bank -> host; `out` (e.g. a page-locked bank) is filled in place"""
        bank = np.zeros(n, abi.site_dtype) if out is None else out
        assert bank.dtype == abi.site_dtype and bank.size >= n and bank.flags.c_contiguous
        self._check(self.lib.omcgpu_get_source_bank(bank.ctypes.data_as(C.c_void_p), C.c_int64(n)))
        return bank[:n]

    def get_fission_bank(self, capacity: int) -> np.ndarray:
        bank = np.zeros(capacity, abi.site_dtype)
        n = C.c_int64(0)
        self._check(self.lib.omcgpu_get_fission_bank(bank.ctypes.data_as(C.c_void_p),
                                                     C.c_int64(capacity), C.byref(n)))
        return bank[:n.value]

    def gen_in(self, n, id_offset, particle_seed_base, keff, tallies_active,
               master_seed=DEFAULT_SEED, stride=DEFAULT_STRIDE, source_bank_clean=False) -> abi.GenIn:
        return abi.GenIn(n, id_offset, particle_seed_base, master_seed, stride, keff,
                         int(tallies_active), int(source_bank_clean))

    def update(self, model: abi.ModelHandle, what: int):
        """omcgpu_update: re-upload cell temperatures / material densities of `model`"""
        self._check(self.lib.omcgpu_update(model.byref(), C.c_uint32(what)))
        self.model = model

    def run_generation(self, gen_in: abi.GenIn) -> abi.GenOut:
        out = abi.GenOut()
        self._check(self.lib.omcgpu_run_generation(C.byref(gen_in), C.byref(out)))
        return out

    def transport_generation(self, gen_in: abi.GenIn, source_bank: np.ndarray,
                             fission_bank: np.ndarray | None = None):
        """Drop-in for transport_event_based() + sort_fission_bank() on host buffers."""
        src = np.ascontiguousarray(source_bank, dtype=abi.site_dtype)
        if fission_bank is None:
            fission_bank = np.zeros(3 * int(gen_in.n_particles), abi.site_dtype)
        out = abi.GenOut()
        self._check(self.lib.omcgpu_transport_generation(
            C.byref(gen_in), src.ctypes.data_as(C.c_void_p),
            fission_bank.ctypes.data_as(C.c_void_p), C.c_int64(fission_bank.size), C.byref(out)))
        return fission_bank[:out.n_fission], out

    def next_generation(self, gen_in: abi.GenIn, source_bank: np.ndarray, seed_id: int) -> abi.GenOut:
        """transport_event_based() + synchronize_bank() of a single-process run on the host
        source bank, which is replaced IN PLACE by the next generation's source."""
        assert source_bank.dtype == abi.site_dtype and source_bank.flags.c_contiguous
        out = abi.GenOut()
        self._check(self.lib.omcgpu_next_generation(
            C.byref(gen_in), source_bank.ctypes.data_as(C.c_void_p), C.c_int64(seed_id), C.byref(out)))
        return out

    def host_register(self, a: np.ndarray):
        self._check(self.lib.omcgpu_host_register(a.ctypes.data_as(C.c_void_p), C.c_int64(a.nbytes)))

    def host_unregister(self, a: np.ndarray):
        self._check(self.lib.omcgpu_host_unregister(a.ctypes.data_as(C.c_void_p)))

    def comb_fission_bank(self, start, total, n_particles, seed_id, master_seed=DEFAULT_SEED,
                          stride=DEFAULT_STRIDE) -> int:
        n = C.c_int64(0)
        self._check(self.lib.omcgpu_comb_fission_bank(
            C.c_int64(start), C.c_int64(total), C.c_int64(n_particles), C.c_int64(seed_id),
            C.c_int64(master_seed), C.c_uint64(stride), C.byref(n)))
        return n.value

    def comb_indices(self, start, n_local, total, n_particles, seed_id, master_seed=DEFAULT_SEED,
                     stride=DEFAULT_STRIDE) -> np.ndarray:
        """host half of the comb (omcgpu_comb_indices): selected sites relative to `start`"""
        idx = np.zeros(int(n_particles) + int(n_local) + 16, np.int32)
        n = C.c_int64(0)
        self._check(self.lib.omcgpu_comb_indices(
            C.c_int64(start), C.c_int64(n_local), C.c_int64(total), C.c_int64(n_particles),
            C.c_int64(seed_id), C.c_int64(master_seed), C.c_uint64(stride), _p(idx, C.c_int32),
            C.c_int64(idx.size), C.byref(n)))
        return idx[:n.value].astype(np.int64)

    def bank_columns_device(self, which: int):
        cols = (C.c_void_p * 12)()
        n = C.c_int64(0)
        cap = C.c_int64(0)
        self._check(self.lib.omcgpu_bank_columns_device(C.c_int(which), cols, C.byref(n), C.byref(cap)))
        return [cols[i] for i in range(12)], n.value, cap.value

    def set_source_count(self, n: int):
        self._check(self.lib.omcgpu_set_source_count(C.c_int64(n)))

    # -- multi-GPU (include/omcgpu.h: omcgpu_comm_*, omcgpu_sync_bank, omcgpu_reduce_tallies) ----
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(abi.OMCGPU_COMM_ID_BYTES)
        self._check(self.lib.omcgpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        assert len(uid) == abi.OMCGPU_COMM_ID_BYTES
        self._check(self.lib.omcgpu_comm_init(C.c_char_p(uid), C.c_int(rank), C.c_int(n_ranks)))

    def comm_init_custom(self, rank, n_ranks, allgather, sendrecv, allreduce):
        """callbacks are ctypes function objects (kept alive here)"""
        self._comm_callbacks = (allgather, sendrecv, allreduce)
        self._check(self.lib.omcgpu_comm_init_custom(C.c_int(rank), C.c_int(n_ranks), allgather, sendrecv,
                                                     allreduce, None))

    def comm_finalize(self):
        self._check(self.lib.omcgpu_comm_finalize())

    def comm_allgather(self, values, n_ranks):
        send = np.ascontiguousarray(values, np.int64)
        recv = np.zeros((n_ranks, send.size), np.int64)
        self._check(self.lib.omcgpu_comm_allgather_i64(_p(send, C.c_int64), C.c_int32(send.size),
                                                       _p(recv, C.c_int64)))
        return recv

    def sync_bank(self, n_particles, work_index, seed_id, globals_, master_seed=DEFAULT_SEED,
                  stride=DEFAULT_STRIDE, host_bank: np.ndarray | None = None):
        """synchronize_bank over the ranks + the sum of `globals_`; -> (reduced globals, n_source)"""
        wi = np.ascontiguousarray(work_index, np.int64)
        g = np.ascontiguousarray(globals_, np.float64).copy()
        n = C.c_int64(0)
        hb = host_bank.ctypes.data_as(C.c_void_p) if host_bank is not None else None
        self._check(self.lib.omcgpu_sync_bank(C.c_int64(n_particles), _p(wi, C.c_int64), C.c_int64(seed_id),
                                              C.c_int64(master_seed), C.c_uint64(stride), _p(g, C.c_double),
                                              C.c_int32(g.size), hb, C.byref(n)))
        return g, n.value

    def reduce_tallies(self, root: int = -1):
        self._check(self.lib.omcgpu_reduce_tallies(C.c_int(root)))

    # -- tallies -----------------------------------------------------------------
    def tally_results(self, tally: int, value: np.ndarray):
        """value += device VALUE array of `tally`; zeroes the device array."""
        assert value.dtype == np.float64 and value.flags.c_contiguous
        self._check(self.lib.omcgpu_tally_results(C.c_int32(tally), _p(value, C.c_double),
                                                  C.c_int64(value.size)))

    # -- probes ------------------------------------------------------------------
    def calculate_xs(self, E, material, max_nuc: int = 0):
        E = np.ascontiguousarray(E, np.float64)
        material = np.ascontiguousarray(material, np.int32)
        n = E.size
        out = np.zeros((n, 4))
        ig = np.zeros((n, max(max_nuc, 1)), np.int32)
        self._check(self.lib.omcgpu_calculate_xs(
            _p(E, C.c_double), _p(material, C.c_int32), C.c_int64(n), _p(out, C.c_double),
            _p(ig, C.c_int32) if max_nuc else None, C.c_int32(max_nuc)))
        return out, ig

    def find_cell(self, xyz, uvw=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        cell = np.zeros(n, np.int32); inst = np.zeros(n, np.int32)
        mat = np.zeros(n, np.int32); nco = np.zeros(n, np.int32)
        up = None
        if uvw is not None:
            uvw = np.ascontiguousarray(uvw, np.float64)
            up = _p(uvw, C.c_double)
        self._check(self.lib.omcgpu_find_cell(_p(xyz, C.c_double), up, C.c_int64(n),
                                              _p(cell, C.c_int32), _p(inst, C.c_int32),
                                              _p(mat, C.c_int32), _p(nco, C.c_int32)))
        return cell, inst, mat, nco

    def distance_to_boundary(self, xyz, uvw):
        xyz = np.ascontiguousarray(xyz, np.float64)
        uvw = np.ascontiguousarray(uvw, np.float64)
        n = xyz.shape[0]
        d = np.zeros(n); s = np.zeros(n, np.int32); lv = np.zeros(n, np.int32)
        lt = np.zeros((n, 3), np.int32)
        self._check(self.lib.omcgpu_distance_to_boundary(
            _p(xyz, C.c_double), _p(uvw, C.c_double), C.c_int64(n), _p(d, C.c_double),
            _p(s, C.c_int32), _p(lv, C.c_int32), _p(lt, C.c_int32)))
        return d, s, lv, lt

    def collide(self, material, E, uvw, seed, keff, wgt):
        """omcgpu_collide: the collision kernels' device function on independent neutrons.
        -> dict(E, u, wgt, k_abs, seed, n_sites, sites[(neutron, progeny) order])"""
        material = np.ascontiguousarray(material, np.int32)
        E = np.ascontiguousarray(E, np.float64)
        uvw = np.ascontiguousarray(uvw, np.float64)
        seed = np.ascontiguousarray(seed, np.uint64)
        n = E.size
        out = dict(E=np.zeros(n), u=np.zeros((n, 3)), wgt=np.zeros(n), k_abs=np.zeros(n),
                   seed=np.zeros(n, np.uint64), n_sites=np.zeros(n, np.int32))
        cap = 3 * self.max_particles
        sites = np.zeros(cap, abi.site_dtype)
        tot = C.c_int64(0)
        self._check(self.lib.omcgpu_collide(
            C.c_int64(n), _p(material, C.c_int32), _p(E, C.c_double), _p(uvw, C.c_double),
            _p(seed, C.c_uint64), C.c_double(keff), C.c_double(wgt), _p(out["E"], C.c_double),
            _p(out["u"], C.c_double), _p(out["wgt"], C.c_double), _p(out["k_abs"], C.c_double),
            _p(out["seed"], C.c_uint64), _p(out["n_sites"], C.c_int32), sites.ctypes.data_as(C.c_void_p),
            C.c_int64(cap), C.byref(tot)))
        out["sites"] = sites[:tot.value]
        return out

    def future_seed(self, n, seed):
        n = np.ascontiguousarray(n, np.uint64); seed = np.ascontiguousarray(seed, np.uint64)
        out = np.zeros_like(n)
        self._check(self.lib.omcgpu_future_seed(_p(n, C.c_uint64), _p(seed, C.c_uint64),
                                                C.c_int64(n.size), _p(out, C.c_uint64)))
        return out

    def prn(self, seed, count):
        seed = np.ascontiguousarray(seed, np.uint64)
        out = np.zeros((seed.size, count))
        self._check(self.lib.omcgpu_prn(_p(seed, C.c_uint64), C.c_int64(seed.size),
                                        C.c_int32(count), _p(out, C.c_double)))
        return out


# ------------------------------------------------------------------------------
# eigenvalue statistics shared by the driver
# ------------------------------------------------------------------------------
def combined_keff(n, gsum, gsq, k_col_abs, k_col_tra, k_abs_tra, keff, keff_std):
    """openmc_get_keff, src/eigenvalue.cpp:420-577."""
    FP_REL_PRECISION = 1e-5
    if n <= 3:
        return keff, (math.inf if n <= 1 else keff_std)
    kv = [gsum[abi.OMCGPU_K_COLLISION] / n, gsum[abi.OMCGPU_K_ABSORPTION] / n,
          gsum[abi.OMCGPU_K_TRACKLENGTH] / n]
    cov = [[0.0] * 3 for _ in range(3)]
    cov[0][0] = (gsq[abi.OMCGPU_K_COLLISION] - n * kv[0] * kv[0]) / (n - 1)
    cov[1][1] = (gsq[abi.OMCGPU_K_ABSORPTION] - n * kv[1] * kv[1]) / (n - 1)
    cov[2][2] = (gsq[abi.OMCGPU_K_TRACKLENGTH] - n * kv[2] * kv[2]) / (n - 1)
    cov[0][1] = cov[1][0] = (k_col_abs - n * kv[0] * kv[1]) / (n - 1)
    cov[0][2] = cov[2][0] = (k_col_tra - n * kv[0] * kv[2]) / (n - 1)
    cov[1][2] = cov[2][1] = (k_abs_tra - n * kv[1] * kv[2]) / (n - 1)

    def same(a, b):
        return (abs(kv[a] - kv[b]) / kv[a] < FP_REL_PRECISION and
                abs(cov[a][a] - cov[b][b]) / cov[a][a] < FP_REL_PRECISION)
    pair = None
    if same(0, 1):
        pair = (0, 2)
    elif same(0, 2):
        pair = (0, 1)
    elif same(1, 2):
        pair = (0, 1)
    if pair is None:
        g = 0.0
        S = [0.0, 0.0, 0.0]
        k0 = 0.0
        for l, (i, j, k) in enumerate(((0, 1, 2), (1, 2, 0), (2, 0, 1))):
            f = (cov[j][j] * (cov[k][k] - cov[i][k]) - cov[k][k] * cov[i][j] +
                 cov[j][k] * (cov[i][j] + cov[i][k] - cov[j][k]))
            S[0] += f * cov[0][l]
            S[1] += (cov[j][j] + cov[k][k] - 2.0 * cov[j][k]) * kv[l] * kv[l]
            S[2] += (cov[k][k] + cov[i][j] - cov[j][k] - cov[i][k]) * kv[l] * kv[j]
            k0 += f * kv[l]
            g += f
        S = [s * (n - 1) for s in S]
        S[0] *= (n - 1) * (n - 1)
        k0 /= g
        g *= (n - 1) * (n - 1)
        return k0, math.sqrt(S[0] / (g * n * (n - 3)) * (1 + n * ((S[1] - 2 * S[2]) / g)))
    i, j = pair
    f = kv[i] - kv[j]
    g = cov[i][i] + cov[j][j] - 2.0 * cov[i][j]
    k0 = kv[i] - (cov[i][i] - cov[i][j]) / g * f
    var = (cov[i][i] * cov[j][j] - cov[i][j] * cov[i][j]) * (g + n * f * f) / (n * (n - 2) * g * g)
    return k0, math.sqrt(var)


def comb_indices(start, n_local, total, n_particles, seed_id, master_seed=DEFAULT_SEED,
                 stride=DEFAULT_STRIDE):
    """Uniform comb of synchronize_bank (src/eigenvalue.cpp:152-178), host side."""
    from .rng import init_seed, prn
    seed = init_seed(seed_id, 0, master_seed, stride)
    teeth_distance = float(total) / n_particles
    xi, seed = prn(seed)
    teeth_offset = xi * teeth_distance
    end = start + n_local
    tooth_start = math.ceil((start - teeth_offset) / teeth_distance)
    tooth_end = math.floor((end - teeth_offset) / teeth_distance) + 1
    count = max(tooth_end - tooth_start, 0)
    # tooth += teeth_distance, repeated: np.cumsum accumulates left to right in
    # double precision, i.e. exactly the reference's sequential recurrence
    steps = np.full(count, teeth_distance)
    if count:
        steps[0] = tooth_start * teeth_distance + teeth_offset
    teeth = np.cumsum(steps)
    return np.floor(teeth).astype(np.int64) - start


class Simulation:
    """Eigenvalue driver: the L4 bookkeeping of libopenmc around the device path."""

    def __init__(self, engine: Engine, model: abi.ModelHandle, n_particles: int, n_batches: int,
                 n_inactive: int, source_bank: np.ndarray, gen_per_batch: int = 1,
                 seed: int = DEFAULT_SEED, stride: int = DEFAULT_STRIDE, device: int = 0,
                 resident: bool = True, comm=None):
        self.engine, self.model = engine, model
        self.n_particles, self.n_batches, self.n_inactive = n_particles, n_batches, n_inactive
        self.gen_per_batch, self.seed, self.stride = gen_per_batch, seed, stride
        self.device, self.resident, self.comm = device, resident, comm
        self.rank = comm.rank if comm else 0
        self.n_procs = comm.size if comm else 1
        # calculate_work, src/simulation.cpp:657-680
        min_work, rem = divmod(n_particles, self.n_procs)
        self.work_index = [0]
        for r in range(self.n_procs):
            self.work_index.append(self.work_index[-1] + min_work + (1 if r < rem else 0))
        self.work = self.work_index[self.rank + 1] - self.work_index[self.rank]
        self._initial_source = np.ascontiguousarray(source_bank, dtype=abi.site_dtype)
        assert self._initial_source.size == self.work, "source bank must hold this rank's share"
        self.initialized = False

    # -- openmc_simulation_init, src/simulation.cpp:75-167 -----------------------
    def simulation_init(self):
        self.engine.init(self.model, self.device, max(self.work, 1))   # work_per_rank, as allocate_banks
        if self.comm is not None and hasattr(self.comm, "attach"):
            self.comm.attach(self.engine)      # the communicator lives on the engine's device
        self.source_bank = self._initial_source.copy()
        self.engine.set_source_bank(self.source_bank)
        self.current_batch = 0
        self.total_gen = 0
        self.keff = 1.0
        self.keff_std = 0.0
        self.k_generation = []
        self.k_sum = [0.0, 0.0]
        self.gt = np.zeros((4, 3))            # global_tallies [VALUE, SUM, SUM_SQ]
        self.k_col_abs = self.k_col_tra = self.k_abs_tra = 0.0
        self.n_realizations = 0
        nt = int(self.model["n_tallies"])
        ofs = self.model["tally_result_ofs"]
        self.tally_results = []
        for t in range(nt):
            ns = int(self.model["tally_n_scores"][t])
            nb = int(ofs[t + 1] - ofs[t]) // ns
            self.tally_results.append(np.zeros((nb, ns, 3)))
        self.tally_n_realizations = 0
        self.stats = dict(ms_transport=0.0, ms_xs=0.0, n_events=0, n_xs_lookups=0, n_collisions=0,
                          n_crossings=0, n_lost=0, generations=0, active_ms=0.0, active_particles=0)
        self.initialized = True

    def overall_generation(self, gen):
        return self.gen_per_batch * (self.current_batch - 1) + gen

    # -- openmc_next_batch, src/simulation.cpp:232-280 ----------------------------
    def next_batch(self):
        assert self.initialized, "Simulation has not been initialized yet."
        self.current_batch += 1
        total_weight = 0.0
        active = self.current_batch > self.n_inactive
        for gen in range(1, self.gen_per_batch + 1):
            og = self.overall_generation(gen)
            keff_generation = self.gt[abi.OMCGPU_K_TRACKLENGTH, 0]     # initialize_generation
            gi = self.engine.gen_in(self.work, self.work_index[self.rank],
                                    (self.total_gen + og - 1) * self.n_particles, self.keff,
                                    active, self.seed, self.stride)
            if self.resident:
                out = self.engine.run_generation(gi)
            else:
                self.fission_bank, out = self.engine.transport_generation(gi, self.source_bank)
            if getattr(self, "on_generation", None):
                self.on_generation(out)
            st = self.stats
            st["ms_transport"] += out.ms_transport; st["ms_xs"] += out.ms_xs
            st["n_events"] += out.n_events; st["n_xs_lookups"] += out.n_xs_lookups
            st["n_collisions"] += out.n_collisions; st["n_crossings"] += out.n_crossings
            st["n_lost"] += out.n_lost; st["generations"] += 1
            if active:
                st["active_ms"] += out.ms_transport; st["active_particles"] += self.work
            g = np.array(out.global_tally[:] + [out.total_weight])
            if self.comm is not None and getattr(self.comm, "native", False):
                # one library call: the comb + exchange of synchronize_bank AND the sum of the globals
                g = self.comm.sync_bank(self, self.total_gen + og, g)
                total_weight += g[4]
                self.gt[:, 0] += g[:4]
            else:
                if self.comm:                    # MPI_Allreduce / MPI_Reduce of globals
                    g = self.comm.allreduce_sum(g)
                total_weight += g[4]
                self.gt[:, 0] += g[:4]           # finalize_generation :524-533
                self._synchronize_bank(og, out.n_fission)
            # calculate_generation_keff, src/eigenvalue.cpp:50-84
            keff_generation = self.gt[abi.OMCGPU_K_TRACKLENGTH, 0] - keff_generation
            self.k_generation.append(keff_generation / self.n_particles)
            self._calculate_average_keff(og, gen, active)
        self._finalize_batch(total_weight, active)

    def _synchronize_bank(self, og, n_fission):
        """synchronize_bank, src/eigenvalue.cpp:86-365."""
        if n_fission == 0:
            raise OmcGpuError(f"No fission sites banked on rank {self.rank}")
        seed_id = self.total_gen + og
        if self.comm is None:
            if self.resident:
                n = self.engine.comb_fission_bank(0, n_fission, self.n_particles, seed_id,
                                                  self.seed, self.stride)
                assert n == self.n_particles, (n, self.n_particles)
            else:
                idx = comb_indices(0, n_fission, n_fission, self.n_particles, seed_id, self.seed,
                                   self.stride)
                self.source_bank = np.ascontiguousarray(self.fission_bank[idx])
            return
        self.comm.synchronize_bank(self, seed_id, n_fission)
        if not self.resident:
            # the exchange leaves the new source in the DEVICE bank; the host bank is what the
            # next generation uploads (transport_generation), so it has to follow
            self.source_bank = np.ascontiguousarray(self.engine.get_source_bank(self.work))

    def _calculate_average_keff(self, og, gen, active):
        """src/eigenvalue.cpp:367-418"""
        i = og - 1
        n = self.gen_per_batch * self.n_realizations + gen if active else 0
        if n <= 0:
            self.keff = self.k_generation[i]
        else:
            self.k_sum[0] += self.k_generation[i]
            self.k_sum[1] += self.k_generation[i] ** 2
            self.keff = self.k_sum[0] / n
            if n > 1:
                v = (self.k_sum[1] / n - self.keff ** 2) / (n - 1)
                self.keff_std = math.sqrt(v) if v >= 0 else 0.0

    def _finalize_batch(self, total_weight, active):
        """finalize_batch -> accumulate_tallies, src/tallies/tally.cpp:1059-1104, 821-878."""
        self.n_realizations += 1
        if active:
            k_col = self.gt[abi.OMCGPU_K_COLLISION, 0] / total_weight
            k_abs = self.gt[abi.OMCGPU_K_ABSORPTION, 0] / total_weight
            k_tra = self.gt[abi.OMCGPU_K_TRACKLENGTH, 0] / total_weight
            self.k_col_abs += k_col * k_abs
            self.k_col_tra += k_col * k_tra
            self.k_abs_tra += k_abs * k_tra
        for i in range(4):
            val = self.gt[i, 0] / total_weight
            self.gt[i, 0] = 0.0
            self.gt[i, 1] += val
            self.gt[i, 2] += val * val
        if active and self.tally_results:
            norm = 1.0 / (float(self.n_particles) * self.gen_per_batch)
            self.tally_n_realizations += 1
            native = self.comm is not None and getattr(self.comm, "native", False)
            if native:
                self.engine.reduce_tallies(-1)             # reduce_tally_results :997-1057, on the device
            for t, res in enumerate(self.tally_results):
                value = np.zeros(res.shape[0] * res.shape[1])
                self.engine.tally_results(t, value)        # hand-off into results_(:,:,VALUE)
                if self.comm and not native:
                    value = self.comm.allreduce_sum(value)  # reduce_tally_results :997-1057
                val = value.reshape(res.shape[:2]) * norm
                res[:, :, 1] += val
                res[:, :, 2] += val * val
        if self.current_batch <= self.n_inactive:            # src/simulation.cpp:416-419
            self.gt[:] = 0.0
            self.n_realizations = 0

    # -- openmc.lib-style conveniences -------------------------------------------
    def run(self):
        if not self.initialized:
            self.simulation_init()
        while self.current_batch < self.n_batches:
            self.next_batch()
        return self.keff_combined()

    def iter_batches(self):
        if not self.initialized:
            self.simulation_init()
        while self.current_batch < self.n_batches:
            self.next_batch()
            yield self.current_batch

    def keff_combined(self):
        return combined_keff(self.n_realizations, self.gt[:, 1], self.gt[:, 2], self.k_col_abs,
                             self.k_col_tra, self.k_abs_tra, self.keff, self.keff_std)

    def global_tallies(self):
        return self.gt.copy()

    def get_source_bank(self):
        if self.resident:
            return self.engine.get_source_bank(self.work)
        return self.source_bank

    def tally_mean_std(self, t):
        """mean and standard deviation of the mean per bin (openmc.lib Tally.mean/std_dev)."""
        res = self.tally_results[t]
        n = self.tally_n_realizations
        mean = res[:, :, 1] / n
        if n > 1:
            var = np.maximum(res[:, :, 2] / n - mean ** 2, 0.0) / (n - 1)
        else:
            var = np.full_like(mean, np.inf)
        return mean, np.sqrt(var)

    def simulation_finalize(self):
        self.total_gen += self.current_batch * self.gen_per_batch   # src/simulation.cpp:192
        self.engine.finalize()
        self.initialized = False
