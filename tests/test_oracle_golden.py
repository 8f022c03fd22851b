"""The oracle (oracle/oracle.cpp) against the reference's own golden results.

tests/testing_harness.py:89-124 of the reference digests a statepoint into
``k-combined`` and every tally ``sum`` / ``sum_sq`` at %12.6E; the strings below
are copied from the cited results_true.dat files.  Matching them digit for
digit pins RNG streams, source sampling, geometry, MG physics, bank sort, the
uniform comb, k statistics and tally accumulation of the restatement."""
import pytest

import cases
from oracle.binding import Oracle


@pytest.mark.parametrize("name", list(cases.GOLDEN))
def test_golden(name):
    model, box, of = cases.build(name)
    o = Oracle(model)
    src = o.sample_initial_source(1000, box, of)
    r = o.run_eigenvalue(src, 10, 5)
    k, s, tallies = cases.GOLDEN[name]
    assert cases.fmt(r["k_combined"][0]) == k
    assert cases.fmt(r["k_combined"][1]) == s
    assert r["n_lost"] == 0
    for i, (ts, tsq) in enumerate(tallies):
        assert cases.fmt(r["tally_sum"][i]) == ts
        assert cases.fmt(r["tally_sum_sq"][i]) == tsq
