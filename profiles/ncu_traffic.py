#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of bench.py's in-loop kernels.

    python profiles/ncu_traffic.py <workload> <report.ncu-rep> <trace file> <generation> <sweep> [digest name]

The capture holds the four launches of ONE sweep of the paired event loop (k_surface_xa,
k_collide_xa and the two k_score launches that score the PREVIOUS sweep's flights), taken with
`-k regex:'k_surface_xa|k_collide_xa|k_score' -s S -c 4`, S computed from the launch list so that it
is the first launch of sweep <sweep> of generation <generation> (profiles/exp_r56.sh).  The engine
writes OMCGPU_TRACE (one line per sweep: generation, sweep, collisions, crossings), so the list
length of every captured launch is known and DRAM bytes are reported PER LIST ITEM; bench.py scales
them to its average launch (`roofline.kernels.*.traffic`).
"""
import csv
import io
import json
import os
import subprocess
import sys


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def num(x):
    return float(x.replace(",", ""))


def main():
    workload, rep, trace, gen, sweep = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    source = sys.argv[6] if len(sys.argv) > 6 else os.path.basename(rep)
    hdr, units, data = rows_of(rep)
    col = {h: i for i, h in enumerate(hdr)}

    def scaled(d, name):       # ncu prints large counters in Kbyte / Mbyte / Gbyte, times in us / ms
        return num(d[col[name]]) * UNIT.get(units[col[name]], 1.0)
    counts = {}
    for ln in open(trace):
        p = ln.split()
        if len(p) >= 4:
            counts[(int(p[0]), int(p[1]))] = (int(p[2]), int(p[3]))     # collisions, crossings
    n_coll, n_surf = counts[(gen, sweep)]
    p_coll, p_surf = counts[(gen, sweep - 1)]
    result = {}
    scores = []
    for d in data:
        name = d[col["Kernel Name"]]
        e = {"grid": int(num(d[col["launch__grid_size"]])),
             "dram_bytes": scaled(d, "dram__bytes_read.sum") + scaled(d, "dram__bytes_write.sum"),
             "duration_us": scaled(d, "gpu__time_duration.sum"),
             "l2_hit_pct": num(d[col["lts__t_sector_hit_rate.pct"]]),
             "l1_hit_pct": num(d[col["l1tex__t_sector_hit_rate.pct"]]),
             "lanes_active": num(d[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
             "dram_pct_of_peak": num(d[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
             "warps_active_pct": num(d[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
             "issue_active_pct": num(d[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
             "long_scoreboard_per_issue": num(
                 d[col["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]]),
             "source": source, "generation": gen, "sweep": sweep}
        if "k_surface_xa" in name:
            e["items"] = n_surf
            result["k_surface_xa"] = e
        elif "k_collide_xa" in name:
            e["items"] = n_coll
            result["k_collide_xa"] = e
        elif "k_score" in name:
            scores.append(e)
    if scores:       # the larger launch scores the crossing list's flights, the other the collision list's
        scores.sort(key=lambda e: -e["duration_us"])
        scores[0]["items"] = p_surf
        if len(scores) > 1:
            scores[1]["items"] = p_coll
        tot = {"dram_bytes": sum(s["dram_bytes"] for s in scores), "items": sum(s["items"] for s in scores),
               "duration_us": sum(s["duration_us"] for s in scores), "l2_hit_pct": scores[0]["l2_hit_pct"],
               "lanes_active": scores[0]["lanes_active"], "source": source, "generation": gen, "sweep": sweep - 1,
               "launches": scores}
        result["k_score"] = tot
    for e in result.values():
        e["dram_bytes_per_item"] = e["dram_bytes"] / max(e["items"], 1)
    json_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    try:
        allj = json.load(open(json_path))
    except Exception:
        allj = {}
    probe = allj.get(workload, {}) if isinstance(allj.get(workload), dict) else {}
    keep = {k: v for k, v in probe.items() if k in ("lookups", "dram_bytes_per_launch", "source")}
    allj[workload] = dict(result, **({"k_probe_xs": keep} if keep else {}))
    allj["_comment"] = ("per in-loop kernel: one mid-generation launch from an `ncu --set full` capture of bench.py "
                        "(profiles/exp_r56.sh); items = list length of that launch from the engine's sweep trace; "
                        "dram_bytes_per_item = (dram__bytes_read.sum + dram__bytes_write.sum) / items")
    json.dump(allj, open(json_path, "w"), indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "launches"} for k, v in result.items()}, indent=1))


if __name__ == "__main__":
    main()
