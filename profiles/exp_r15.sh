#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r15; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d xs_frac=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], r.get("frac")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -2 | python -c "$summ" ) | tee -a $O/matrix.log; }
run default X=1
run l2persist OMCGPU_L2_PERSIST=1
run fetch64 OMCGPU_L2_FETCH=64
run fetch128 OMCGPU_L2_FETCH=128
run cg OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/cg.so
run cg_persist OMCGPU_L2_PERSIST=1 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/cg.so
