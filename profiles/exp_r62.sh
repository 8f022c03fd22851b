#!/bin/bash
# r62 (1 GPU): full GPU tier + smoke at HEAD (rotational_wedge moved to the statistical tier)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r62; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/smoke.log
