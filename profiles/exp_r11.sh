#!/bin/bash
# experiment matrix 3: order-preserving list build (index-sorted lists)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r11; mkdir -p $O
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
run ordered OMCGPU_ORDERED=1
run ordered_again OMCGPU_ORDERED=1
run unordered OMCGPU_ORDERED=0
run unordered_sweep1 OMCGPU_SWEEP=1
run ordered_prefetch OMCGPU_PREFETCH=1
run ordered_priority OMCGPU_PRIORITY=1
run ordered_part4 OMCGPU_SCORE_PART=4
run ordered_part16 OMCGPU_SCORE_PART=16
run ordered_tail4k OMCGPU_TAIL_THRESHOLD=4096
run ordered_tail64k OMCGPU_TAIL_THRESHOLD=65536
for v in x3c3 x4c2 x4c3; do run ordered_$v OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/$v.so; done
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "== C3 assembly_ce" | tee -a $O/matrix.log; timeout 400 $B --workload assembly_ce 2>&1 | tail -1 | tee $O/bench_C3.json | python -c "$summ" | tee -a $O/matrix.log
echo "== C4 c5g7_core" | tee -a $O/matrix.log; timeout 400 $B --workload c5g7_core 2>&1 | tail -1 | tee $O/bench_C4.json | python -c "$summ" | tee -a $O/matrix.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_list.log 2>&1
# active generation = 2nd: skip the warm-up generation's ~500 event kernels, capture one sweep's kernels mid-way
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s 620 -c 4 \
   -o $O/sweep_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
ls -la $O
