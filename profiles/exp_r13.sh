#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r13; mkdir -p $O
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
run default X=1
run s3 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/s3.so
run s1 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/s1.so
run part8 OMCGPU_SCORE_PART=8
run part32 OMCGPU_SCORE_PART=32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_score' -s 320 -c 2 \
   -o $O/score_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
