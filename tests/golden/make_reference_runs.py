"""Writes tests/golden/refrun_<case>.npz: what the reference's OWN binary (oracle/_ref/openmc =
the unmodified sources of /root/reference compiled against oracle/h5shim, see oracle/Makefile)
computes END TO END for the cases of tests/reference_runs.py -- k per generation, k-combined,
global tallies, every tally's sum / sum_sq, the final source bank.  The GPU box has no
/root/reference but carries these files; there the CUDA path is held to them
(tests/test_gpu_reference_runs.py), and on either box the oracle is (CPU tier).

    python tests/golden/make_reference_runs.py        (in the build container)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import reference_runs  # noqa: E402
from oracle import binding  # noqa: E402

if __name__ == "__main__":
    binding.build()
    for name in reference_runs.RUNS:
        d = reference_runs.run_reference(name, threads=1)
        np.savez_compressed(reference_runs.GOLDEN % name, **d)
        print(name, "k_combined", d["k_combined"], "bins", d["tally_sum"].size)
