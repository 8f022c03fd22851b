#!/bin/bash
# r77 (1 GPU): multigroup density multipliers (extended build): full GPU tier + smoke + C2 / C4 sanity
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
ROOT=$PWD
O=$ROOT/gpurun_out/r77; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/smoke.log
for w in "--particles 4000000" "--particles 2000000 --workload c5g7_core"; do
python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 3 --warmup 3 $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms %.1f k=%.5f'%(d['value'], d['ms_per_step'], d['k_combined'][0]))"
done
