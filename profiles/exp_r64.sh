#!/bin/bash
# r64 (1 GPU): ncu --set full of two mid-generation k_score2 launches (timed generation, counters off)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r64; mkdir -p $O
CMD="python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_score' -s 80 -c 2 -o $O/score2_full -f $CMD > $O/ncu_full.log 2>&1
python profiles/ncu_metrics.py $O/score2_full.ncu-rep > $O/score2_ncu.txt 2>&1
tail -3 $O/ncu_full.log
