#!/bin/bash
# round-end profile set of the default configuration
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r22; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/gpu_tests.log
timeout 900 python bench.py 2>&1 | tail -1 > $O/bench_default.json; cat $O/bench_default.json | cut -c1-400
timeout 600 python bench.py --particles 1000000 --no-cpu-baseline 2>&1 | tail -1 > $O/bench_1e6.json; cut -c1-300 $O/bench_1e6.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s 620 -c 4 \
   -o $O/sweep_full -f python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_xs -s 4 -c 1 \
   -o $O/probe_xs_full -f python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_probe.log 2>&1
ls -la $O
