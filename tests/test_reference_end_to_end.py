"""END TO END against the reference itself: oracle/_ref/openmc is every translation unit of
libopenmc + src/main.cpp, unmodified, compiled from /root/reference against oracle/h5shim
(recipe: oracle/Makefile).  It reads a model.xml and the data libraries this repo writes
(oracle/refrun.py) and runs whole eigenvalue calculations.

* the binary reproduces the reference's own regression goldens (results_true.dat), which
  validates the shim and the writers;
* the oracle reproduces, BIT FOR BIT, what the binary computes for the CE and MG
  configurations of BASELINE.json: k of every generation, k-combined, global tallies, every
  tally bin's sum and sum of squares (tracklength and collision estimators; mesh, energy,
  cell, material, distribcell filters), the final source bank;
* the reference's --event mode gives the same numbers as its history mode;
* tests/golden/refrun_*.npz (what the GPU tier is held to) are what the binary writes today.
"""
import os
import shutil
import tempfile

import numpy as np
import pytest

import cases
import reference_runs as rr
from oracle import refrun
from oracle.binding import Oracle, build

HAVE_REF = os.path.isdir("/root/reference/src")


@pytest.fixture(scope="module")
def reference_binary():
    if HAVE_REF:
        build()
    if not refrun.available():
        pytest.skip("oracle/_ref/openmc not built (no /root/reference on this box)")
    return refrun.BINARY




def _mg_case(name):
    """description + library of the reference's MG regression cases (tests/cases.py)"""
    from openmc_b200 import examples
    if name == "mg_legendre":
        m, box, of = examples.slab_mg()
        return m, examples.slab_2g_library(1), box, of
    if name == "mg_max_order":
        m, box, of = examples.slab_mg()
        m.settings["max_order"] = 1
        return m, examples.slab_2g_library(2), box, of
    if name == "mg_survival_biasing":
        m, box, of = examples.slab_mg()
        m.settings["survival_biasing"] = 1
        return m, examples.slab_2g_library(1), box, of
    if name == "mg_basic":
        return examples.mg_basic()
    return examples.mg_temperature_multi()


@pytest.mark.parametrize("name", list(cases.GOLDEN))
def test_binary_reproduces_the_regression_goldens(reference_binary, name):
    """tests/regression_tests/<name>/results_true.dat digit for digit, from the binary built
    against oracle/h5shim reading the mgxs.h5 / model.xml written by oracle/refrun.py"""
    m, lib, box, of = _mg_case(name)
    d = tempfile.mkdtemp(prefix="omc_gold_")
    try:
        rr.write_inputs(d, m, lib, box, of, False, 1000, 10, 5)
        r = refrun.run(d, threads=2)
        assert r["returncode"] == 0, r["stdout"][-1500:] + r["stderr"][-1500:]
        g = rr.digest(r)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    k, s, tallies = cases.GOLDEN[name]
    assert cases.fmt(g["k_combined"][0]) == k
    assert cases.fmt(g["k_combined"][1]) == s
    for i, (ts, tsq) in enumerate(tallies):
        assert cases.fmt(g["tally_sum"][i]) == ts
        assert cases.fmt(g["tally_sum_sq"][i]) == tsq


def _oracle_run(name):
    m, lib, model, box, of, ce = rr.build(name)
    particles, batches, inactive = rr.RUNS[name]
    o = Oracle(model)
    src = o.sample_initial_source(particles, box, of)
    return o.run_eigenvalue(src, batches, inactive)


def _assert_oracle_equals(r, g):
    assert r["n_lost"] == 0
    assert np.array_equal(r["k_generation"], g["k_generation"])
    assert np.array_equal(np.array(r["k_combined"]), g["k_combined"])
    assert np.array_equal(r["global_sum"], g["global_tallies"][:, 0])
    assert np.array_equal(r["global_sum_sq"], g["global_tallies"][:, 1])
    assert r["tally_sum"].size == g["tally_sum"].size
    assert np.array_equal(r["tally_sum"], g["tally_sum"])
    assert np.array_equal(r["tally_sum_sq"], g["tally_sum_sq"])
    b = r["source_bank"]
    assert np.array_equal(b["r"], g["bank_r"]) and np.array_equal(b["u"], g["bank_u"])
    assert np.array_equal(b["E"], g["bank_E"]) and np.array_equal(b["wgt"], g["bank_wgt"])
    assert np.array_equal(b["delayed_group"], g["bank_delayed_group"])


@pytest.mark.parametrize("name", list(rr.RUNS))
def test_oracle_equals_the_reference_end_to_end(reference_binary, name):
    _assert_oracle_equals(_oracle_run(name), rr.run_reference(name, threads=1))


@pytest.mark.parametrize("name", list(rr.RUNS))
def test_oracle_equals_the_committed_reference_runs(name):
    """the same bar on a box without the reference: the committed files the binary wrote"""
    g = dict(np.load(rr.GOLDEN % name))
    _assert_oracle_equals(_oracle_run(name), g)


@pytest.mark.parametrize("name", ["pincell_ce", "c5g7_core"])
def test_committed_reference_runs_are_current(reference_binary, name):
    g = dict(np.load(rr.GOLDEN % name))
    now = rr.run_reference(name, threads=1)
    for k in g:
        assert np.array_equal(g[k], now[k]), k


@pytest.mark.parametrize("name", ["pincell_ce", "assembly_ce", "c5g7_core"])
def test_reference_event_mode_equals_history_mode(reference_binary, name):
    """the reference's own CI re-runs its suite with --event: same k, same banks;
    tallies to rounding (the event loop scores in a different order)"""
    h = rr.run_reference(name, threads=1)
    e = rr.run_reference(name, threads=1, event=True)
    assert np.array_equal(h["k_generation"], e["k_generation"])
    assert np.array_equal(h["bank_E"], e["bank_E"])
    np.testing.assert_allclose(e["tally_sum"], h["tally_sum"], rtol=1e-11, atol=1e-300)


def test_reference_threads_do_not_change_the_answer(reference_binary):
    a = rr.run_reference("assembly_ce", threads=1)
    b = rr.run_reference("assembly_ce", threads=4)
    # per-thread partial sums: equal to rounding, and the same fission sites
    np.testing.assert_allclose(b["k_generation"], a["k_generation"], rtol=1e-13)
    assert np.array_equal(a["bank_E"], b["bank_E"])
    np.testing.assert_allclose(b["tally_sum"], a["tally_sum"], rtol=1e-11, atol=1e-300)
