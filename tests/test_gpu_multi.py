"""GPU tier, N > 1: the library's own NCCL path (omcgpu_comm_init, omcgpu_sync_bank,
omcgpu_reduce_tallies) on two GPUs reproduces the one-GPU run -- particle ids and seeds are global
and every rank combs with the same seed, so the redistributed source bank is identical site by
site and k / tallies differ only by the order of the floating-point sums.  Skipped on a box with
one GPU (the CPU tier runs the same library code over gloo callbacks, tests/test_multi_rank.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(size, case, tmp_path, tag):
    port = _free_port()
    procs = []
    for r in range(size):
        out = str(tmp_path / f"{tag}_rank{r}.npz")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "nccl_worker.py"), str(r), str(size),
                                       str(port), case, out]))
    for p in procs:
        assert p.wait(timeout=600) == 0
    return [np.load(tmp_path / f"{tag}_rank{r}.npz") for r in range(size)]


@pytest.mark.parametrize("case", ["mg_temperature_multi"])
def test_two_gpus_over_nccl_equal_one_gpu(case, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    single = _launch(1, case, tmp_path, "one")[0]
    ranks = _launch(2, case, tmp_path, "two")
    for r in ranks:
        np.testing.assert_allclose(r["kgen"], single["kgen"], rtol=1e-12)
    np.testing.assert_allclose(ranks[0]["k"], single["k"], rtol=1e-9)
    bank = np.concatenate([r["bank"] for r in ranks])
    assert int(ranks[1]["lo"]) == ranks[0]["bank"].size
    for f in ("r", "u", "E", "wgt", "delayed_group", "parent_id", "progeny_id"):
        assert np.array_equal(bank[f], single["bank"][f]), f
    for i in range(2):
        np.testing.assert_allclose(ranks[0][f"t{i}"], single[f"t{i}"], rtol=1e-10, atol=1e-300)
