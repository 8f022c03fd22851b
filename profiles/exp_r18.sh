#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r18; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"]))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run ool X=1
run prev OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/prev.so
run ool_again X=1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --particles 100000"
run ool_1e5 X=1
run prev_1e5 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/prev.so
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload assembly_ce"
run ool_C3 X=1
run prev_C3 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/prev.so
