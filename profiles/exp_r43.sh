#!/bin/bash
# r43 (1 GPU): new mesh DDA / scoring kernel vs r42's base; default bench with sub-lines
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r43; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
run() { name=$1; shift
  ( env "$@" timeout 400 $B $EXTRA 2>$O/$name.err | tail -1 > $O/$name.json )
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
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/gpu_tests.log
EXTRA="--particles 1000000"; run base_1e6 A=1
run dynscore_1e6 OMCGPU_DYNAMIC_SCORE=1
run part8_1e6 OMCGPU_SCORE_PART=8
run part32_1e6 OMCGPU_SCORE_PART=32
EXTRA="--particles 4000000"; run base_4e6 A=1
EXTRA="--particles 2000000 --workload pincell_mg"; run c1 A=1
EXTRA="--particles 2000000 --workload assembly_ce"; run c3 A=1
EXTRA="--particles 2000000 --workload c5g7_core"; run c4 A=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_score' -s 110 -c 2 \
   -o $O/score_full -f python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $O/ncu_score.log 2>&1
ls -la $O | head -30
