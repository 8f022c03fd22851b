"""Worker of tests/test_gpu_multi.py: one rank of an N-GPU eigenvalue run through the library's
native NCCL path (omcgpu_comm_init / omcgpu_sync_bank / omcgpu_reduce_tallies).

    python tests/nccl_worker.py <rank> <size> <port> <case> <out.npz>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, size, port, case, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    import torch.distributed as dist
    import cases
    from openmc_b200 import lib
    from openmc_b200.comm import NativeComm
    from oracle.binding import Oracle
    N, BATCHES, INACTIVE = 20000, 5, 2
    model, box, of = cases.build(case)
    full = Oracle(model).sample_initial_source(N, box, of)
    comm = None
    if size > 1:
        store = dist.TCPStore("127.0.0.1", port, size, is_master=(rank == 0))
        comm = NativeComm.over_nccl(rank, size, store)
    lo = sum(N // size + (1 if r < N % size else 0) for r in range(rank))
    hi = lo + N // size + (1 if rank < N % size else 0)
    sim = lib.Simulation(lib.Engine(), model, N, BATCHES, INACTIVE, full[lo:hi], device=rank, comm=comm)
    k = sim.run()
    bank = sim.get_source_bank()
    np.savez(out, k=np.array(k), kgen=np.array(sim.k_generation), bank=bank, lo=lo,
             **{f"t{i}": t[:, :, 1].copy() for i, t in enumerate(sim.tally_results)})
    sim.simulation_finalize()


if __name__ == "__main__":
    main()
