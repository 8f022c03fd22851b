"""GPU tier: the reference's own binary with the hooks compiled in (oracle/_ref/openmc_gpu, built in
the development container from the unmodified sources + openmc_b200/host) running its transport
on the B200 through libomcgpu.so: `OPENMC_GPU=1 openmc_gpu` in a directory with model.xml and the
data library, statepoint read back -- the user-visible drop-in.  Held to the end-to-end runs of
the unmodified binary (tests/golden/refrun_*.npz) with the bars of test_gpu_reference_runs.py."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import reference_runs as rr
from oracle import omh5

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BINARY = os.path.join(ROOT, "oracle", "_ref", "openmc_gpu")
ENGINE = os.path.join(ROOT, "openmc_b200", "csrc", "libomcgpu.so")


def _run(name, gpu_engine):
    if not os.path.exists(GPU_BINARY):
        pytest.skip("oracle/_ref/openmc_gpu was not built in the development container")
    m, lib, model, box, of, ce = rr.build(name)
    particles, batches, inactive = rr.RUNS[name]
    d = tempfile.mkdtemp(prefix="omc_gpu_dropin_")
    try:
        rr.write_inputs(d, m, lib, box, of, ce, particles, batches, inactive)
        env = dict(os.environ, OPENMC_GPU="1", OMCGPU_LIB=ENGINE)
        res = subprocess.run([GPU_BINARY, "-s", "2"], cwd=d, capture_output=True, text=True, env=env, timeout=900)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        sp = sorted(f for f in os.listdir(d) if f.startswith("statepoint"))[-1]
        return rr.digest({"statepoint": omh5.File(os.path.join(d, sp))}), batches - inactive
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("name", ["pincell_mg", "c5g7_core", "rotated_fills"])
def test_dropin_multigroup(gpu_engine, name):
    got, _ = _run(name, gpu_engine)
    ref = dict(np.load(rr.GOLDEN % name))
    np.testing.assert_allclose(got["k_generation"], ref["k_generation"], rtol=1e-6)
    np.testing.assert_allclose(got["k_combined"], ref["k_combined"], rtol=2e-5)
    if ref["tally_sum"].size:
        np.testing.assert_allclose(got["tally_sum"], ref["tally_sum"], rtol=1e-6, atol=1e-9 * ref["tally_sum"].max())


@pytest.mark.parametrize("name", ["pincell_ce", "pincell_ce_temperatures", "assembly_ce"])
def test_dropin_continuous_energy(gpu_engine, name):
    got, n = _run(name, gpu_engine)
    ref = dict(np.load(rr.GOLDEN % name))
    assert abs(got["k_generation"][0] - ref["k_generation"][0]) < 2e-3 * ref["k_generation"][0]
    sig = np.hypot(got["k_combined"][1], ref["k_combined"][1])
    assert abs(got["k_combined"][0] - ref["k_combined"][0]) < 3 * sig
    mean_g, mean_r = got["tally_sum"] / n, ref["tally_sum"] / n
    var_g = np.maximum(got["tally_sum_sq"] / n - mean_g ** 2, 0) / (n - 1)
    var_r = np.maximum(ref["tally_sum_sq"] / n - mean_r ** 2, 0) / (n - 1)
    ok = (mean_r > 0) & (var_r + var_g > 0)
    z = (mean_g[ok] - mean_r[ok]) / np.sqrt(var_g[ok] + var_r[ok])
    assert np.all(np.abs(z) < 5.0), np.abs(z).max()
    assert np.mean(z ** 2) < 2.0
