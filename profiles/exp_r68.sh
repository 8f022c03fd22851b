#!/bin/bash
# r68-r69 (1 GPU): after record_lost stopped taking PState by reference: the build against variants without the
# hexagonal-lattice dispatch (-DOMC_NO_HEX) and without the lost-particle records (-DOMC_NO_LOST_REC)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r69; mkdir -p $O
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
for w in "--particles 4000000" "--particles 2000000 --workload assembly_ce" "--particles 2000000 --workload c5g7_core"; do
EXTRA="$w"
run head A=1
run no_hex OMCGPU_LIB=$V/no_hex.so

done
