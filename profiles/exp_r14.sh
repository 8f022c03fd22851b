#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r14; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d xs_frac=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], r.get("frac")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.log
run replicas8 X=1
run replicas1 OMCGPU_TALLY_REPLICAS=1
run replicas32 OMCGPU_TALLY_REPLICAS=32
run batch_all_4 OMCGPU_BATCH_XS=1 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/ball.so
run batch_all_2 OMCGPU_BATCH_XS=1 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/ball3.so
run sync16 OMCGPU_SYNC_EVERY=16
run sync4 OMCGPU_SYNC_EVERY=4
