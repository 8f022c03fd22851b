#!/bin/bash
# r49 (re-run of r48 after the device-side bank sync in the drop-in) (1 GPU): the drop-in binary (reference + hooks) on the B200; default bench with the drop-in leg
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r49; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -8 | tee $O/gpu_dropin_tests.log
timeout 900 python bench.py --no-other-configs 2>$O/bench.err | tail -1 > $O/bench.json
python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4e e2e %.4e ms %.1f dropin %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], json.dumps(d['dropin'])[:700])); print(json.dumps(d['cpu_baseline'])[:500])"
tail -3 $O/bench.err
