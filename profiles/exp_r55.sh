#!/bin/bash
# r55 (1 GPU): full GPU tier + default bench after white BC, quadrics and the generic-base geometry access
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r55; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/smoke.log
timeout 1200 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json
python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4e e2e %.4e ms %.1f dropin %.4e'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['dropin']['value'])); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d['other_configs'].items()}); print(d['dropin'])"
tail -3 $O/bench.err
