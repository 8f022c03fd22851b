#!/bin/bash
# r42: persistent work-fetching kernels, keyed (class-sorted) lists, block prefix sums for C5
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r42; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
run() { # name, env..., (EXTRA = bench args)
  name=$1; shift
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
for E in "OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1" "OMCGPU_KEYED=1" "OMCGPU_KEYED=1 OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1"; do
  echo "== tests with $E"
  env $E timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_geometry.py -m gpu -x -q 2>&1 | tail -3
done | tee $O/gpu_tests_variants.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pwr_core or many_nuclide" 2>&1 | tail -2 | tee $O/gpu_tests_c5.log
for P in 1000000 4000000; do
EXTRA="--particles $P"
run base_$P A=1
run dyn_$P OMCGPU_DYNAMIC=1
run dynscore_$P OMCGPU_DYNAMIC_SCORE=1
run dynboth_$P OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1
run keyed_$P OMCGPU_KEYED=1
run keyed_dyn_$P OMCGPU_KEYED=1 OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1
done
EXTRA="--particles 100000"
run base_1e5 A=1
run keyed_dyn_1e5 OMCGPU_KEYED=1 OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1
EXTRA="--particles 2000000 --workload assembly_ce"
run c3_base A=1
run c3_keyed_dyn OMCGPU_KEYED=1 OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1
EXTRA="--particles 2000000 --workload c5g7_core"
run c4_base A=1
run c4_keyed_dyn OMCGPU_KEYED=1 OMCGPU_DYNAMIC=1 OMCGPU_DYNAMIC_SCORE=1
EXTRA="--particles 400000 --workload pwr_core_ce"
run c5_noprefix_4e5 OMCGPU_XS_PREFIX=0
run c5_prefix_4e5 A=1
EXTRA="--particles 1000000 --workload pwr_core_ce"
run c5_prefix_1e6 A=1
EXTRA="--particles 2000000 --workload pwr_core_ce"
run c5_prefix_2e6 A=1
