#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r17; mkdir -p $O
# latency chain of one event: the event kernels with the GPU nearly empty (30 000 histories)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa' -s 40 -c 2 \
   -o $O/lowcount_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --particles 30000 > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log | cut -c1-300
