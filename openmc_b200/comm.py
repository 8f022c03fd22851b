"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL over
NVLink/NVSwitch on the box, gloo in the CPU tests).

The path shards the way the reference's MPI decomposition does:
every rank owns the global source sites [work_index[r], work_index[r+1]) of a
generation, geometry and data are replicated, and there is ONE exchange step
per generation -- ``synchronize_bank`` (src/eigenvalue.cpp:86-365):

1. all-gather of the per-rank fission-bank sizes -> start offsets and total
   (replaces MPI_Exscan + MPI_Bcast, :100-112);
2. every rank combs its (device-resident, sorted) fission bank with the SAME
   generation seed (:152-178) -- omcgpu_comb_fission_bank;
3. all-gather of the sampled counts (replaces MPI_Exscan + MPI_Allgather, :186-197);
4. point-to-point exchange of contiguous runs of sites to their owners
   (replaces MPI_Isend/Irecv, :224-348) -- batched isend/irecv of packed SoA rows
   straight from / into the device bank columns;
5. all-reduce of the k / leakage / weight globals and of tally VALUE arrays
   (replaces MPI_Allreduce/MPI_Reduce, src/eigenvalue.cpp:63, src/tallies/tally.cpp:1014-1053).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

try:  # only TorchComm / GlooTransport (host-side plumbing of the CPU tests) need torch
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = dist = None

_COL_DTYPES = ([torch.float64] * 9 + [torch.int32, torch.int64, torch.int64]) if torch else []


class _CudaPtr:
    """Expose a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


def _column_tensor(ptr, n, dtype, on_gpu):
    if n == 0:
        return torch.zeros(0, dtype=dtype, device="cuda" if on_gpu else "cpu")
    if on_gpu:
        ts = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
        return torch.as_tensor(_CudaPtr(ptr, n, ts), device="cuda")
    ct = {torch.float64: C.c_double, torch.int32: C.c_int32, torch.int64: C.c_int64}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
    return torch.from_numpy(arr)


class TorchComm:
    def __init__(self, on_gpu: bool):
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.on_gpu = on_gpu
        self.dev = "cuda" if on_gpu else "cpu"
        self.exchange_bytes = 0
        self._connect_all()

    def _connect_all(self):
        """One tiny send/recv with every peer: NCCL sets a point-to-point connection up on first
        use (tens to hundreds of ms per pair), and which neighbours a rank trades sites with
        changes from generation to generation -- without this a timed generation can pay for a
        connection (first 8-GPU run: 190 ms per generation outside the device loop)."""
        if self.size == 1:
            return
        ops, keep = [], []
        for peer in range(self.size):
            if peer == self.rank:
                continue
            s = torch.zeros(1, dtype=torch.float64, device=self.dev)
            r = torch.zeros(1, dtype=torch.float64, device=self.dev)
            keep += [s, r]
            ops.append(dist.P2POp(dist.isend, s, peer))
            ops.append(dist.P2POp(dist.irecv, r, peer))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if self.on_gpu:
            torch.cuda.synchronize()

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allgather_int(self, v: int):
        t = torch.tensor([v], dtype=torch.int64, device=self.dev)
        out = [torch.zeros_like(t) for _ in range(self.size)]
        dist.all_gather(out, t)
        return [int(x.item()) for x in out]

    def _bank(self, engine, which, n):
        cols, _, cap = engine.bank_columns_device(which)
        return [_column_tensor(p, cap, dt, self.on_gpu) for p, dt in zip(cols, _COL_DTYPES)]

    def synchronize_bank(self, sim, seed_id, n_fission):
        eng = sim.engine
        counts = self.allgather_int(n_fission)
        total = sum(counts)
        start = sum(counts[:self.rank])
        n_s = eng.comb_fission_bank(start, total, sim.n_particles, seed_id, sim.seed, sim.stride)
        sampled = self.allgather_int(n_s)
        assert sum(sampled) == sim.n_particles, (sampled, sim.n_particles)
        pos = np.concatenate([[0], np.cumsum(sampled)])        # bank_position
        wi = sim.work_index
        cols = self._bank(eng, 0, n_s)
        my_lo, my_hi = int(pos[self.rank]), int(pos[self.rank + 1])
        need_lo, need_hi = wi[self.rank], wi[self.rank + 1]
        n_new = need_hi - need_lo

        def pack(lo, hi):
            # rows [lo, hi) (global site numbers) of my sampled sites as 12 x 8 bytes (ints
            # widened / bit-cast): only the rows that change owner are packed
            return torch.stack([c[lo - my_lo:hi - my_lo] if c.dtype == torch.float64 else
                                c[lo - my_lo:hi - my_lo].to(torch.int64).view(torch.float64)
                                for c in cols], dim=1).contiguous()

        ops, recvs = [], []
        for peer in range(self.size):
            if peer == self.rank:
                continue
            lo, hi = max(my_lo, wi[peer]), min(my_hi, wi[peer + 1])          # what I hold that `peer` owns
            if hi > lo:
                chunk = pack(lo, hi)
                ops.append(dist.P2POp(dist.isend, chunk, peer))
                self.exchange_bytes += chunk.numel() * 8
            lo, hi = max(int(pos[peer]), need_lo), min(int(pos[peer + 1]), need_hi)   # what `peer` holds that I own
            if hi > lo:
                buf = torch.empty((hi - lo, 12), dtype=torch.float64, device=self.dev)
                recvs.append((lo - need_lo, hi - need_lo, buf))
                ops.append(dist.P2POp(dist.irecv, buf, peer))
        # the sites that stay: one contiguous shift per column (send chunks were packed above)
        keep_lo, keep_hi = max(my_lo, need_lo), min(my_hi, need_hi)
        if keep_hi > keep_lo and keep_lo - my_lo != keep_lo - need_lo:
            for c in cols:
                c[keep_lo - need_lo:keep_hi - need_lo] = c[keep_lo - my_lo:keep_hi - my_lo].clone()
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for a, b, buf in recvs:
            for k, c in enumerate(cols):
                if c.dtype == torch.float64:
                    c[a:b] = buf[:, k]
                else:
                    c[a:b] = buf[:, k].contiguous().view(torch.int64).to(c.dtype)
        if self.on_gpu:
            torch.cuda.synchronize()
        eng.set_source_count(n_new)


class NativeComm:
    """The product's N > 1 path: everything inter-rank happens INSIDE libomcgpu.so
    (omcgpu_sync_bank / omcgpu_reduce_tallies over NCCL, include/omcgpu.h); this class only
    carries rank / size and the one-off distribution of the NCCL unique id.  The communicator
    lives on the engine's device, so it joins in ``attach`` -- which ``Simulation.simulation_init``
    calls right after ``omcgpu_init``.

    ``NativeComm.over_nccl(rank, size, store)``: rank 0 makes the id, `store` (anything with
    set / get, e.g. torch.distributed.TCPStore -- a plain TCP key-value store, no process group)
    hands it to the others.  ``NativeComm.over_callbacks(transport)`` plugs a caller-supplied
    transport (GlooTransport below in the CPU tests)."""
    native = True

    def __init__(self, rank, size, connect):
        self.rank, self.size, self._connect, self.engine = rank, size, connect, None

    @classmethod
    def over_nccl(cls, rank, size, store, key="omcgpu_nccl_id"):
        joined = [0]      # a communicator per engine: every attach publishes its id under its own key

        def connect(engine):
            k = "%s_%d" % (key, joined[0])
            joined[0] += 1
            if rank == 0:
                uid = engine.comm_unique_id()
                store.set(k, uid)
            else:
                uid = bytes(store.get(k))
            engine.comm_init(uid, rank, size)
        return cls(rank, size, connect)

    @classmethod
    def over_callbacks(cls, transport):
        def connect(engine):
            engine.comm_init_custom(transport.rank, transport.size, transport.allgather_cb,
                                    transport.sendrecv_cb, transport.allreduce_cb)
        c = cls(transport.rank, transport.size, connect)
        c.transport = transport
        return c

    def attach(self, engine):
        if self.engine is not engine:
            self._connect(engine)
            self.engine = engine

    def sync_bank(self, sim, seed_id, globals_, host_bank=None):
        g, n = self.engine.sync_bank(sim.n_particles, sim.work_index, seed_id, globals_, sim.seed,
                                     sim.stride, host_bank)
        assert n == sim.work, (n, sim.work)
        return g

    def barrier(self):
        self.engine.comm_allgather([0], self.size)

    def max_over_ranks(self, x: float) -> float:
        """(integers on the wire: nanoseconds for a time)"""
        return float(self.engine.comm_allgather([int(round(x * 1e9))], self.size).max()) * 1e-9

    def sum_over_ranks(self, x: int) -> int:
        return int(self.engine.comm_allgather([int(x)], self.size).sum())


class GlooTransport:
    """omcgpu_comm_init_custom callbacks over torch.distributed (gloo) for the CPU test tier,
    where "device" buffers of the host-compiled engine are host memory: the library's own
    synchronize_bank code (csrc/comm.h) then runs across processes without a GPU."""

    def __init__(self):
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int64))
        SR = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                         C.POINTER(C.c_void_p), C.POINTER(C.c_int64))
        AR = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_int32)
        self.allgather_cb = AG(self._allgather)
        self.sendrecv_cb = SR(self._sendrecv)
        self.allreduce_cb = AR(self._allreduce)

    def _allgather(self, user, send, n, recv):
        try:
            t = torch.tensor([send[k] for k in range(n)], dtype=torch.int64)
            out = [torch.zeros_like(t) for _ in range(self.size)]
            dist.all_gather(out, t)
            for r in range(self.size):
                for k in range(n):
                    recv[r * n + k] = int(out[r][k])
            return 0
        except Exception:
            return 1

    def _sendrecv(self, user, n_ops, peer, is_send, buf, nbytes):
        try:
            ops, keep = [], []
            for k in range(n_ops):
                if nbytes[k] == 0:
                    continue
                arr = np.ctypeslib.as_array(C.cast(buf[k], C.POINTER(C.c_uint8)), shape=(nbytes[k],))
                t = torch.from_numpy(arr)
                keep.append(t)
                ops.append(dist.P2POp(dist.isend if is_send[k] else dist.irecv, t, int(peer[k])))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            return 0
        except Exception:
            return 1

    def _allreduce(self, user, buf, n, root):
        try:
            arr = np.ctypeslib.as_array(buf, shape=(n,))
            t = torch.from_numpy(arr)
            if root < 0:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            else:
                dist.reduce(t, root, op=dist.ReduceOp.SUM)
            return 0
        except Exception:
            return 1
