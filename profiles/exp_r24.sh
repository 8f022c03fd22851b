#!/bin/bash
# round-end numbers of the other configurations (parity-test cases, not bench lines)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r24; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f ev/p=%.1f k=%.5f+-%.5f lost=%d e2e=%s xs_frac=%s nuclide_lookups/s=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["events_per_particle"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("nuclide_lookups_per_s")))
    except Exception as e: print("   FAIL", l[:300])
'
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --particles 2000000"
for w in pincell_mg assembly_ce c5g7_core; do
  echo "== $w (2e6 histories)" | tee -a $O/matrix.log; timeout 600 $B --workload $w 2>&1 | tail -1 | tee $O/bench_$w.json | python -c "$summ" | tee -a $O/matrix.log
done
echo "== pwr_core_ce 300 nuclides (4e5 histories)" | tee -a $O/matrix.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload pwr_core_ce --particles 400000 2>&1 | tail -1 | tee $O/bench_pwr_core_ce.json | python -c "$summ" | tee -a $O/matrix.log
