#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r26; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}; il = r.get("in_loop") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f k=%.5f+-%.5f lost=%d xs_ms=%s e2e=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], il.get("ms_total"), (d.get("e2e") or {}).get("value")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 600 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/gpu_tests.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload pwr_core_ce --particles 400000"
run C5_default_sorted_ordered X=1
run C5_sorted_unordered OMCGPU_ORDERED=0
run C5_paired_unsorted OMCGPU_SORT_XS=0
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --particles 1000000"
run C2_default X=1
run C2_sweep1_ordered OMCGPU_SWEEP=1
