#!/bin/bash
# r74 (1 GPU): round-end default bench, both arms, as the driver runs them
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r74; mkdir -p $O
nproc | tee $O/nproc.txt
( time timeout 1500 python bench.py --impl reference 2>$O/bench_reference.err | tail -1 > $O/bench_reference.json ) 2>&1 | grep real
( time timeout 1500 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json ) 2>&1 | grep real
python -c "
import json; d=json.load(open('$O/bench.json')); r=json.load(open('$O/bench_reference.json'))
print('value %.4e e2e %.4e ms %.1f dropin %.4e | reference arm %.4e (%s, %d cores)'%(d['value'], d['e2e']['value'], d['ms_per_step'], (d.get('dropin') or {}).get('value', float('nan')), r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores']))
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'alg', d['roofline']['algorithmic_bytes_per_launch'])
print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items()})
print('cpu_baseline', d['cpu_baseline'])"
tail -3 $O/bench.err
