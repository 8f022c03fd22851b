#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r28; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}; il = r.get("in_loop") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f k=%.5f+-%.5f lost=%d e2e=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], (d.get("e2e") or {}).get("value")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 600 $B 2>&1 | tail -1 | tee $O/bench_$name.json | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload pwr_core_ce --particles 400000"
run C5_sort_coll X=1
run C5_no_sort_coll OMCGPU_SORT_COLL=0
timeout 900 python bench.py 2>&1 | tail -1 > $O/bench_default.json; cut -c1-330 $O/bench_default.json
