#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r20; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 > $O/bench_default.json; cat $O/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json; cat $O/bench_reference.json
