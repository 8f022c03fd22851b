#!/bin/bash
# r46 (1 GPU): after moving the collision-estimator scoring into its own kernel; new GPU tests held to
# the reference binary's end-to-end runs; both bench arms with the reference binary as the CPU side
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r46; mkdir -p $O
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_reference.err | tail -1 > $O/bench_reference.json
timeout 900 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json
python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4e e2e %.4e ms %.1f frac %.3f cpu %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['cpu_baseline'])[:600]))
d=json.load(open('$O/bench_reference.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['history'], d['cpu_baseline']['event'])"
