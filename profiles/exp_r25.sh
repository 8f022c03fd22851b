#!/bin/bash
# energy-sorted lookup on the many-nuclide stress case (C5) and on C2 / C3
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r25; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}; il = r.get("in_loop") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f k=%.5f+-%.5f lost=%d xs_ms=%s lookups=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], il.get("ms_total"), il.get("lookups")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 600 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload pwr_core_ce --particles 400000"
run C5_default X=1
run C5_sweep1_unfused OMCGPU_SWEEP=1 OMCGPU_FUSE_XS=0
run C5_sorted OMCGPU_SORT_XS=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload assembly_ce --particles 1000000"
run C3_default X=1
run C3_sorted OMCGPU_SORT_XS=1
