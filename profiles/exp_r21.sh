#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r21; mkdir -p $O
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
run default_4e6 X=1
run tail32k OMCGPU_TAIL_THRESHOLD=32768
run tail8k OMCGPU_TAIL_THRESHOLD=8192
run x4c4 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/x4c4.so
run s3 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/s3.so
