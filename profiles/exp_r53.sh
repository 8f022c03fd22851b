#!/bin/bash
# r53 (1 GPU): is staging the CSG tables worth anything when the access costs nothing extra?
#   geo_never  = gld() is a plain __ldg (the r51 build)          OMCGPU_GEO_SMEM=0
#   geo_always = gld() is an LDS at a precomputed offset (measurement build: no global fallback)
#   default    = gld() picks shared / global per address (the committed build)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r53; mkdir -p $O
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
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
}
V=$PWD/openmc_b200/csrc/variants
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
for rep in 1 2; do
EXTRA="--particles 4000000"
run never_4e6_$rep OMCGPU_LIB=$V/geo_never.so OMCGPU_GEO_SMEM=0
run always_4e6_$rep OMCGPU_LIB=$V/geo_always.so
run default_4e6_$rep A=1
done
EXTRA="--particles 2000000 --workload assembly_ce"
run c3_never OMCGPU_LIB=$V/geo_never.so OMCGPU_GEO_SMEM=0
run c3_always OMCGPU_LIB=$V/geo_always.so
EXTRA="--particles 2000000 --workload c5g7_core"
run c4_never OMCGPU_LIB=$V/geo_never.so OMCGPU_GEO_SMEM=0
run c4_always OMCGPU_LIB=$V/geo_always.so
EXTRA="--particles 2000000 --workload pincell_mg"
run c1_never OMCGPU_LIB=$V/geo_never.so OMCGPU_GEO_SMEM=0
run c1_always OMCGPU_LIB=$V/geo_always.so
