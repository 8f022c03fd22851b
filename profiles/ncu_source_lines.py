#!/usr/bin/env python
"""Top CUDA source lines of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
    python profiles/ncu_source_lines.py report.ncu-rep k_score [N]
columns: share of warp instructions, average active lanes, share of stall samples."""
import csv
import io
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines = []
    fname = func = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and kernel in (func or ""):
            d = dict(zip(hdr, r))
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(d["Instructions Executed"]),
                              int(d["Thread Instructions Executed"]), int(d["# Samples"])))
            except (KeyError, ValueError):
                pass
    ti = sum(l[3] for l in lines) or 1
    ts = sum(l[5] for l in lines) or 1
    print(f"{kernel}: {ti} warp instructions, {ts} samples")
    for f, n, src, wi, thr, smp in sorted(lines, key=lambda l: -l[5])[:top]:
        print(f"{smp / ts:6.3f} smp {wi / ti:6.3f} inst lanes={thr / max(wi, 1):5.1f}  {f}:{n}: {src[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
