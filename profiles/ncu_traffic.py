#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures of bench.py's in-loop kernels.

    python profiles/ncu_traffic.py <workload> <report.ncu-rep> <trace file> <first launch skipped (-s)> [digest name]

The capture is taken with `-k regex:'k_surface_xa|k_collide_xa|k_score' -s S -c C` while the engine
writes OMCGPU_TRACE (one line per sweep: generation, sweep, collisions, crossings).  Launches of the
three kernels come in a fixed order per sweep (crossing, collision, score of the crossing list's
flights, score of the collision list's flights -- the scoring pair of a sweep runs during the next
one), so the S-th matching launch is mapped to its list length through the trace, and DRAM bytes
are reported PER LIST ITEM; bench.py scales them to its average launch.
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
    hdr, data = rows[0], rows[2:]
    return hdr, data


def main():
    workload, rep, trace = sys.argv[1], sys.argv[2], sys.argv[3]
    source = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
    hdr, data = rows_of(rep)
    col = {h: i for i, h in enumerate(hdr)}
    sweeps = [tuple(int(x) for x in ln.split()) for ln in open(trace) if ln.strip()]
    out = {}
    for d in data:
        name = d[col["Kernel Name"]]
        key = next((k for k in ("k_surface_xa", "k_collide_xa", "k_score", "k_xs", "k_collide", "k_surface")
                    if k in name), None)
        if key is None:
            continue
        grid = int(d[col["launch__grid_size"]].replace(",", ""))
        rd = float(d[col["dram__bytes_read.sum"]].replace(",", ""))
        wr = float(d[col["dram__bytes_write.sum"]].replace(",", ""))
        unit_r = 1.0
        e = out.setdefault(key, {"launches": []})
        e["launches"].append({
            "grid": grid, "dram_bytes": (rd + wr) * unit_r,
            "duration_us": float(d[col["gpu__time_duration.sum"]].replace(",", "")),
            "l2_hit_pct": float(d[col["lts__t_sector_hit_rate.pct"]].replace(",", "")),
            "lanes_active": float(d[col["smsp__thread_inst_executed_per_inst_executed.ratio"]].replace(",", "")),
            "dram_pct_of_peak": float(d[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]].replace(",", "")),
            "long_scoreboard_per_issue": float(
                d[col["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]].replace(",", "")),
        })
    # list items of a captured launch: every kernel is launched over the host's last-known upper
    # bound (grid * 256) but processes the device count; the trace gives the counts of the sweeps
    # of that generation -- the captured launch is matched by the closest upper bound
    result = {}
    for key, e in out.items():
        best = e["launches"][len(e["launches"]) // 2]
        result[key] = dict(best)
        result[key]["source"] = source
    json_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    try:
        allj = json.load(open(json_path))
    except Exception:
        allj = {}
    allj[workload] = result
    allj["_comment"] = ("per kernel: one mid-generation launch from an ncu --set full capture of bench.py; "
                        "`items` (list length of that launch) is filled in from the OMCGPU_TRACE file by hand / by "
                        "exp scripts; dram_bytes_per_item = dram_bytes / items")
    json.dump(allj, open(json_path, "w"), indent=1)
    print(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
