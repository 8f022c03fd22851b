#!/bin/bash
# round-1 experiment matrix (one gpurun call): sweep modes, persistent k_score, tail threshold, build variants
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r09; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; echo "== $name" | tee -a $O/matrix.log; ( env "$@" timeout 300 $B 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('   value=%.4g ms/step=%.1f dev_ms=%.1f launches=%d k=%s' % (d['value'], d['ms_per_step'], d['device_ms_per_step'], d['gpu_launches'], d['k_combined']))
    except Exception as e: print('   FAIL', l[:300])
" ) | tee -a $O/matrix.log; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.log
run sweep1 OMCGPU_SWEEP=1
run sweep2 OMCGPU_SWEEP=2
run sweep2_score1 OMCGPU_SWEEP=2 OMCGPU_SCORE_CTAS=1
run sweep2_score2 OMCGPU_SWEEP=2 OMCGPU_SCORE_CTAS=2
run sweep2_score4 OMCGPU_SWEEP=2 OMCGPU_SCORE_CTAS=4
run sweep1_score2 OMCGPU_SWEEP=1 OMCGPU_SCORE_CTAS=2
run sweep2_tail64k OMCGPU_SWEEP=2 OMCGPU_TAIL_THRESHOLD=65536
run sweep2_tail256k OMCGPU_SWEEP=2 OMCGPU_TAIL_THRESHOLD=262144
run sweep2_sync16 OMCGPU_SWEEP=2 OMCGPU_SYNC_EVERY=16
run sweep2_nooverlap OMCGPU_SWEEP=2 OMCGPU_OVERLAP=0
run sweep2_mb3 OMCGPU_SWEEP=2 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/mb3.so
run sweep2_b128 OMCGPU_SWEEP=2 OMCGPU_LIB=$PWD/openmc_b200/csrc/variants/b128.so
# full bench line of the default configuration (e2e + cpu baseline)
timeout 600 python bench.py 2>&1 | tail -1 > $O/bench_default.json; cat $O/bench_default.json
# ncu: launch list of one bench run + full-set capture of the sweep kernels mid-generation
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s 150 -c 6 \
   -o $O/sweep_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_xs -s 4 -c 1 \
   -o $O/probe_xs_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_probe.log 2>&1
ls -la $O
