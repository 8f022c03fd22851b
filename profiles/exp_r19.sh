#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r19; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"]))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --particles 1000000"
run low150k X=1
run low0 OMCGPU_LOW_COUNT=0
run low50k OMCGPU_LOW_COUNT=50000
run low400k OMCGPU_LOW_COUNT=400000
run low150k_tail64k OMCGPU_TAIL_THRESHOLD=65536
run low150k_tail32k OMCGPU_TAIL_THRESHOLD=32768
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --particles 100000"
run 1e5_low150k X=1
run 1e5_low0 OMCGPU_LOW_COUNT=0
