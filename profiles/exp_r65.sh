#!/bin/bash
# r65 (1 GPU): k_score2 with the part loops out of line (80 registers, 3 blocks / SM) and with 128 registers
# (2 blocks / SM) against the first k_score, C2 at 4e6 histories
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r65; mkdir -p $O
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
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
V=$ROOT/openmc_b200/csrc/variants
EXTRA="--particles 4000000"
run v1 OMCGPU_SCORE_VERSION=1
run v2_p6 A=1
run v2_p12 OMCGPU_SCORE_PART=12
run v2_p16 OMCGPU_SCORE_PART=16
run mb2_p6 OMCGPU_LIB=$V/score_mb2.so
run mb2_p12 OMCGPU_LIB=$V/score_mb2.so OMCGPU_SCORE_PART=12
run mb2_v1 OMCGPU_LIB=$V/score_mb2.so OMCGPU_SCORE_VERSION=1
