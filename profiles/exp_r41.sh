#!/bin/bash
# r41: A/B of the local-memory staging (tail kernel always; crossing / collision kernels by env) and tail thresholds
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r41; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
run() { # name, env..., -- args
  name=$1; shift
  ( env "$@" timeout 300 $B $EXTRA 2>/dev/null | tail -1 > $O/$name.json )
  python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    k=d["roofline"]["kernels"]
    print("%-28s value %.3e ms %.1f dev %.1f | surf %.1f coll %.1f score %.1f (ms/step)"%(sys.argv[2], d["value"], d["ms_per_step"], d["device_ms_per_step"],
        k["k_surface_xa"]["ms_total"]/d["steps"], k["k_collide_xa"]["ms_total"]/d["steps"], k["k_score"]["ms_total"]/d["steps"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.log
OMCGPU_LOCAL_STATE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests_local1.log
OMCGPU_LOCAL_STATE=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ce_generation or goldens or c5g7_generation" 2>&1 | tail -3 | tee $O/gpu_tests_local2.log
for P in 1000000 100000; do
EXTRA="--particles $P"
run base_$P A=1
run local1_$P OMCGPU_LOCAL_STATE=1
run local2_$P OMCGPU_LOCAL_STATE=2
run tail64k_$P OMCGPU_TAIL_THRESHOLD=65536
run tail150k_$P OMCGPU_TAIL_THRESHOLD=150000
run tail4k_$P OMCGPU_TAIL_THRESHOLD=4096
run sync4_$P OMCGPU_SYNC_EVERY=4
run sync16_$P OMCGPU_SYNC_EVERY=16
done
EXTRA="--particles 4000000"
run base_4e6 A=1
run local1_4e6 OMCGPU_LOCAL_STATE=1
EXTRA="--particles 2000000 --workload assembly_ce"
run c3_base A=1
run c3_local1 OMCGPU_LOCAL_STATE=1
