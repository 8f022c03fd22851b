#!/bin/bash
# matrix 4: k_score (scores templated, exact march out of line, occupancy), lockstep lookup on C5, final bench + profiles
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r12; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
summ='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d.get("roofline") or {}
        print("   value=%.4g ms/step=%.1f dev_ms=%.1f ev/p=%.1f launches=%d k=%.5f+-%.5f lost=%d e2e=%s xs_frac=%s lookups/s=%s" % (d["value"], d["ms_per_step"], d["device_ms_per_step"], d["events_per_particle"], d["gpu_launches"], d["k_combined"][0], d["k_combined"][1], d["n_lost"], (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("lookups_per_s")))
    except Exception as e: print("   FAIL", l[:300])
'
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 400 $B 2>&1 | tail -1 | python -c "$summ" ) | tee -a $O/matrix.log; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/gpu_tests.log
run default X=1
run s3 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/s3.so
run x3c2 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/x3c2.so
run part16 OMCGPU_SCORE_PART=16
run s3_part16 OMCGPU_SCORE_PART=16 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/s3.so
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload pwr_core_ce --particles 200000"
run C5_batch X=1
run C5_nobatch OMCGPU_BATCH_XS=0
timeout 600 python bench.py 2>&1 | tail -1 > $O/bench_default.json; cat $O/bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s 620 -c 4 \
   -o $O/sweep_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_xs -s 4 -c 1 \
   -o $O/probe_xs_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_probe.log 2>&1
ls -la $O
