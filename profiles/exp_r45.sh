#!/bin/bash
# r45 (1 GPU): where did 392 -> 481 ms per step come from?  HEAD with r43's flags, the default
# line's legs one by one, and the tree of commit 192befe (gpurun_tmp/old) with the same flags.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r45; mkdir -p $O
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
EXTRA="--particles 4000000"; run head_flags A=1
B="python bench.py --no-cpu-baseline --no-other-configs --steps 3 --warmup 3"
run head_e2e A=1
B="python bench.py --no-other-configs --no-e2e --steps 3 --warmup 3"
run head_cpu A=1
cd $ROOT/gpurun_tmp/old
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
run old_flags A=1
