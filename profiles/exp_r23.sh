#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r23; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"]))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or ce_generation or statistics" 2>&1 | tail -3 | tee $O/gpu_tests.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run emit_4e6 X=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --particles 1000000"
run emit_1e6 X=1
