#!/bin/bash
# experiment matrix 2: sector prefetch, stream priorities, fine hash, register caps; the other configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r10; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f ev/p=%.1f launches=%d k=%.5f+-%.5f lost=%d e2e=%s xs_frac=%s lookups/s=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["events_per_particle"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("lookups_per_s")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/gpu_tests.log
run default OMCGPU_SWEEP=2
run sweep1 OMCGPU_SWEEP=1
run noprefetch OMCGPU_PREFETCH=0
run noprefetch_sweep1 OMCGPU_PREFETCH=0 OMCGPU_SWEEP=1
run nopriority OMCGPU_PRIORITY=0
run fine1 OMCGPU_FINE_PER_LOG_BIN=1
run fine4 OMCGPU_FINE_PER_LOG_BIN=4
run fine64 OMCGPU_FINE_PER_LOG_BIN=64
for v in x2c2 x3c3 x4c2 x4c3; do run $v OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/$v.so; done
run x4c2_sweep1 OMCGPU_SWEEP=1 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/x4c2.so
# the other configs (parity-test cases; throughput recorded for the survey's table)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "== C1 pincell_mg" | tee -a $O/matrix.log; timeout 400 $B --workload pincell_mg 2>&1 | tail -1 | tee $O/bench_C1.json | python -c "$summ" | tee -a $O/matrix.log
echo "== C3 assembly_ce" | tee -a $O/matrix.log; timeout 400 $B --workload assembly_ce 2>&1 | tail -1 | tee $O/bench_C3.json | python -c "$summ" | tee -a $O/matrix.log
echo "== C4 c5g7_core" | tee -a $O/matrix.log; timeout 400 $B --workload c5g7_core 2>&1 | tail -1 | tee $O/bench_C4.json | python -c "$summ" | tee -a $O/matrix.log
echo "== C5 pwr_core_ce 300 nuclides" | tee -a $O/matrix.log; timeout 600 $B --workload pwr_core_ce --particles 200000 --steps 2 2>&1 | tail -1 | tee $O/bench_C5.json | python -c "$summ" | tee -a $O/matrix.log
ls -la $O
