#!/bin/bash
# round 2, first GPU call: new parity tests, the bench line with per-kernel rooflines, launch list,
# full captures of the in-loop kernels (C2) and of the C5 lookup / collision kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r40; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 1200 python bench.py 2>$O/bench_default.err | tail -1 > $O/bench_default.json; cut -c1-600 $O/bench_default.json
timeout 600 python bench.py --particles 1000000 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 > $O/bench_1e6.json; cut -c1-300 $O/bench_1e6.json
timeout 600 python bench.py --particles 100000 --no-cpu-baseline --no-other-configs --no-e2e 2>/dev/null | tail -1 > $O/bench_1e5.json; cut -c1-300 $O/bench_1e5.json
# launch list (serialised, cold): shares per kernel
OMCGPU_TRACE=$O/trace_list.txt timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $O/ncu_list.log 2>&1
# full capture, mid-generation sweeps of the 2nd generation
OMCGPU_TRACE=$O/trace_full.txt timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s 620 -c 4 \
   -o $O/sweep_full -f python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $O/ncu_full.log 2>&1
# C5: lookup kernel + collide + advance in the merged sweep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_xs_b|k_collide_b|k_advance' -s 60 -c 3 \
   -o $O/c5_full -f python bench.py --workload pwr_core_ce --particles 400000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $O/ncu_c5.log 2>&1
ls -la $O
