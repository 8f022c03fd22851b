#!/bin/bash
# r79 (1 GPU): round-end evidence: ncu of the final default build (after r76 moved the rare features out): launch list of one bench command and
# `--set full` of the four launches of one mid-generation sweep of the TIMED generation (generation 1, sweep 40;
# generation 2 of this command is bench.py's counting generation), digests + DRAM bytes per item
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r79; mkdir -p $O
CMD="python bench.py --particles 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs"
OMCGPU_TRACE=$O/trace_list.txt timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
# index (among the three in-loop kernels) of the first launch of generation 1, sweep 40
S=$(python - "$O/launches.csv" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 6]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = rows[hdr].index("Kernel Name")
names = [r[col] for r in rows[hdr + 1:]]
gen = 0
idx = 0          # matching launches seen
sweep_first = None
seen_surface_in_gen = 0
for n in names:
    if "k_init" in n:
        gen += 1
        seen_surface_in_gen = 0
    match = ("k_surface_xa" in n) or ("k_collide_xa" in n) or ("k_score" in n)
    if gen == 1 and "k_surface_xa" in n:
        if seen_surface_in_gen == 40:
            print(idx)
            break
        seen_surface_in_gen += 1
    if match:
        idx += 1
PY
)
echo "skip $S" | tee $O/skip.txt
OMCGPU_TRACE=$O/trace_full.txt timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_collide_xa|k_surface_xa|k_score' -s $S -c 4 \
   -o $O/sweep_full -f $CMD > $O/ncu_full.log 2>&1
python profiles/ncu_metrics.py $O/sweep_full.ncu-rep > $O/sweep_ncu.txt 2>&1
cp profiles/traffic.json $O/traffic_before.json
python profiles/ncu_traffic.py pincell_ce $O/sweep_full.ncu-rep $O/trace_full.txt 1 40 profiles/r79_sweep_ncu.txt | tee $O/traffic.log
cp profiles/traffic.json $O/traffic.json
python profiles/summarize_launches.py $O/launches.csv > $O/launch_summary.txt 2>&1 | tail -3
ls -la $O
# the default bench, both arms, with the traffic of this capture
cp $O/traffic.json profiles/traffic.json
( time timeout 1500 python bench.py --impl reference 2>$O/bench_reference.err | tail -1 > $O/bench_reference.json ) 2>&1 | grep real
( time timeout 1500 python bench.py 2>$O/bench.err | tail -1 > $O/bench.json ) 2>&1 | grep real
python -c "
import json; d=json.load(open('$O/bench.json')); r=json.load(open('$O/bench_reference.json'))
print('value %.4e e2e %.4e ms %.1f dropin %.4e | reference arm %.4e (%s, %d cores)'%(d['value'], d['e2e']['value'], d['ms_per_step'], (d.get('dropin') or {}).get('value', float('nan')), r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores']))
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'alg', d['roofline']['algorithmic_bytes_per_launch'])
print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items()})"
