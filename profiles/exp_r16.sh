#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r16; mkdir -p $O
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%.5f+-%.5f lost=%d e2e=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], (d.get("e2e") or {}).get("value")))
    except Exception as e: print("   FAIL", l[:300])
'
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -25 | tee $O/gpu_tests_full_size.log
for n in 100000 300000 2000000 4000000; do
  echo "== particles $n" | tee -a $O/matrix.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --particles $n 2>&1 | tail -1 | python -c "$summ" | tee -a $O/matrix.log
done
