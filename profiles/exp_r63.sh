#!/bin/bash
# r63 (1 GPU): k_score2 (per-flight ray constants, short parts, lanes in step) against the first k_score:
# tally / reference-run GPU tests, then A/B of the step at several part sizes
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r63; mkdir -p $O
run() { name=$1; shift
  ( env "$@" timeout 600 $B $EXTRA 2>$O/$name.err | tail -1 > $O/$name.json )
  python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    k=d["roofline"]["kernels"]
    ms=lambda n: (k[n]["ms_total"]/d["steps"]) if n in k else 0.0
    names=[n for n in ("k_surface_xa","k_collide_xa","k_surface","k_collide","k_xs","k_score") if n in k]
    print("%-26s value %.3e ms %.1f dev %.1f | "%(sys.argv[2], d["value"], d["ms_per_step"], d["device_ms_per_step"]) + " ".join("%s %.1f"%(n[2:],ms(n)) for n in names) + " | k=%.5f lost=%d"%(d["k_combined"][0], d["n_lost"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_reference_runs.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
EXTRA="--particles 4000000"
run v1 OMCGPU_SCORE_VERSION=1
run v2_p6 A=1
run v2_p4 OMCGPU_SCORE_PART=4
run v2_p8 OMCGPU_SCORE_PART=8
run v2_p12 OMCGPU_SCORE_PART=12
EXTRA="--particles 2000000 --workload assembly_ce"
run c3_v1 OMCGPU_SCORE_VERSION=1
run c3_v2 A=1
EXTRA="--particles 2000000 --workload c5g7_core"
run c4_v1 OMCGPU_SCORE_VERSION=1
run c4_v2 A=1
