#!/bin/bash
# r52 (1 GPU): CSG tables staged in shared memory (gld / stage_geometry): GPU tests, A/B against OMCGPU_GEO_SMEM=0
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r52; mkdir -p $O
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
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3"
EXTRA="--particles 4000000"; run smem_4e6 A=1
run nosmem_4e6 OMCGPU_GEO_SMEM=0
EXTRA="--particles 1000000"; run smem_1e6 A=1
run nosmem_1e6 OMCGPU_GEO_SMEM=0
EXTRA="--particles 2000000 --workload assembly_ce"; run c3_smem A=1
run c3_nosmem OMCGPU_GEO_SMEM=0
EXTRA="--particles 2000000 --workload c5g7_core"; run c4_smem A=1
run c4_nosmem OMCGPU_GEO_SMEM=0
EXTRA="--particles 2000000 --workload pincell_mg"; run c1_smem A=1
run c1_nosmem OMCGPU_GEO_SMEM=0
EXTRA="--particles 2000000 --workload pwr_core_ce"; run c5_smem A=1
run c5_nosmem OMCGPU_GEO_SMEM=0
