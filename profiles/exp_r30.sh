#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r30; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -m gpu -q -k "next_generation or goldens or pwr_core" 2>&1 | tail -3 | tee $O/gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 > $O/bench_default.json; cut -c1-300 $O/bench_default.json
