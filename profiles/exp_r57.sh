#!/bin/bash
# r57 (1 GPU): same-box A/B of the current build against the tree of commit 32ca049 (r46: before the
# temperature tables, white BC, quadrics and the shared-memory geometry access)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r57; mkdir -p $O
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
for rep in 1 2; do
cd $ROOT
EXTRA="--particles 4000000"; run head_4e6_$rep A=1
run never_4e6_$rep OMCGPU_LIB=$V/geo_never.so OMCGPU_GEO_SMEM=0
EXTRA="--particles 2000000 --workload assembly_ce"; run head_c3_$rep A=1
cd $ROOT/gpurun_tmp/old46
EXTRA="--particles 4000000"; run old46_4e6_$rep A=1
EXTRA="--particles 2000000 --workload assembly_ce"; run old46_c3_$rep A=1
done
