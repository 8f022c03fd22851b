#!/bin/bash
# r44 (1 GPU): state of HEAD at session start: GPU tests, smoke, default bench line, reference arm
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r44; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json
python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4e e2e %.4e ms %.1f frac %.3f cpu %.3e'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value']))"
