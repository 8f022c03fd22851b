"""N > 1 path on the CPU tier: world_size-2 `gloo`, the engine sources
host-compiled (tests/_emu), through the same TorchComm / Simulation code the
GPU bench uses with NCCL.  The bank exchange must reproduce the single-rank run:
particle ids and seeds are global (src/simulation.cpp:592, 613-616) and every
rank combs with the same seed (src/eigenvalue.cpp:152-178), so the source bank
after synchronize_bank is identical and k differs only by fp reduction order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N, BATCHES, INACTIVE = 1502, 5, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(name, emu_path, comm, resident=True):
    import cases
    from openmc_b200 import lib
    from oracle.binding import Oracle
    model, box, of = cases.build(name)
    full = Oracle(model).sample_initial_source(N, box, of)
    if comm == "native":
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
    else:
        rank, size = (comm.rank, comm.size) if comm else (0, 1)
    # this rank's share of the global source bank (calculate_work, src/simulation.cpp:657-680)
    lo = sum(N // size + (1 if r < N % size else 0) for r in range(rank))
    hi = lo + N // size + (1 if rank < N % size else 0)
    eng = lib.Engine(emu_path)
    if comm == "native":
        # the library's own synchronize_bank (csrc/comm.h) over a caller-supplied transport
        from openmc_b200.comm import GlooTransport, NativeComm
        comm = NativeComm.over_callbacks(GlooTransport())
    sim = lib.Simulation(eng, model, N, BATCHES, INACTIVE, full[lo:hi], comm=comm,
                         resident=resident)
    k = sim.run()
    bank = sim.get_source_bank()
    tallies = [t[:, :, 1].copy() for t in sim.tally_results]
    kgen = list(sim.k_generation)
    sim.simulation_finalize()
    return dict(k=k, kgen=kgen, bank=bank, lo=lo, tallies=tallies)


def _worker(rank, size, port, name, emu_path, out_dir, resident=True, native=False):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from openmc_b200.comm import TorchComm
    r = _run(name, emu_path, "native" if native else TorchComm(on_gpu=False), resident)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), k=np.array(r["k"]), kgen=np.array(r["kgen"]),
             bank=r["bank"], lo=r["lo"], **{f"t{i}": t for i, t in enumerate(r["tallies"])})
    dist.destroy_process_group()


@pytest.mark.parametrize("name,size", [("mg_legendre", 2), ("mg_temperature_multi", 2), ("mg_legendre", 4)])
def test_ranks_equal_one_rank(emu_lib, name, size, tmp_path):
    """world sizes 2 and 4 (N = 1500 is not divisible by 4: uneven shares, and a rank's sampled
    sites go to several owners)"""
    single = _run(name, emu_lib, None)
    port = _free_port()
    mp.spawn(_worker, args=(size, port, name, emu_lib, str(tmp_path)), nprocs=size, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(size)]
    r0 = ranks[0]
    for r in ranks:
        np.testing.assert_allclose(r["kgen"], single["kgen"], rtol=1e-12)
    np.testing.assert_allclose(r0["k"], single["k"], rtol=1e-9)
    # the redistributed source bank: rank r holds global sites [work_index[r], work_index[r+1])
    bank = np.concatenate([r["bank"] for r in ranks])
    assert int(ranks[1]["lo"]) == r0["bank"].size
    for f in ("r", "u", "E", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(bank[f], single["bank"][f]), f
    for i, t in enumerate(single["tallies"]):
        np.testing.assert_allclose(r0[f"t{i}"], t, rtol=1e-10, atol=1e-300)


def test_host_bank_path_with_ranks(emu_lib, tmp_path):
    """Simulation(resident=False) uploads the HOST source bank every generation: with ranks the
    exchanged bank has to come back to the host, or every generation restarts from the initial
    source (round-1 advisor finding)"""
    single = _run("mg_legendre", emu_lib, None)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, "mg_legendre", emu_lib, str(tmp_path), False), nprocs=2, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    for r in ranks:
        np.testing.assert_allclose(r["kgen"], single["kgen"], rtol=1e-12)
    bank = np.concatenate([r["bank"] for r in ranks])
    for f in ("r", "u", "E", "parent_id", "progeny_id"):
        assert np.array_equal(bank[f], single["bank"][f]), f


@pytest.mark.parametrize("name,size", [("mg_legendre", 2), ("mg_temperature_multi", 2), ("mg_legendre", 4)])
def test_native_sync_bank_equals_one_rank(emu_lib, name, size, tmp_path):
    """the same bar for the product's path: omcgpu_sync_bank / omcgpu_reduce_tallies inside the
    library (csrc/comm.h), here over omcgpu_comm_init_custom with gloo callbacks"""
    single = _run(name, emu_lib, None)
    port = _free_port()
    mp.spawn(_worker, args=(size, port, name, emu_lib, str(tmp_path), True, True), nprocs=size, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(size)]
    for r in ranks:
        np.testing.assert_allclose(r["kgen"], single["kgen"], rtol=1e-12)
    np.testing.assert_allclose(ranks[0]["k"], single["k"], rtol=1e-9)
    bank = np.concatenate([r["bank"] for r in ranks])
    for f in ("r", "u", "E", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(bank[f], single["bank"][f]), f
    for i, t in enumerate(single["tallies"]):
        np.testing.assert_allclose(ranks[0][f"t{i}"], t, rtol=1e-10, atol=1e-300)
