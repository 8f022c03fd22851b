#!/bin/bash
# r76 (1 GPU): axis cones, quadrics, white / rotational periodic boundaries and the analog records moved to the
# extended build too: full GPU tier, smoke, C2 (twice) / C3 / C4 on the default build
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r76; mkdir -p $O
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
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/smoke.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
EXTRA="--particles 4000000"
run c2 A=1
run c2_again A=1
run c2_extlib OMCGPU_LIB=$ROOT/openmc_b200/csrc/libomcgpu_ext.so
EXTRA="--particles 2000000 --workload assembly_ce"
run c3 A=1
EXTRA="--particles 2000000 --workload c5g7_core"
run c4 A=1
