#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel totals/shares, split per generation (k_init marks a generation)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    gen = -1
    per = collections.OrderedDict()
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("omc::", "")
        if name.startswith("void cub"):
            name = "cub::" + re.sub(r"<.*", "", name.split("::")[1])
        v = float(row["Metric Value"].replace(",", "")) / 1e6
        if name == "k_init":
            gen += 1
        if gen < 0:
            continue
        d = per.setdefault(gen, collections.OrderedDict())
        d.setdefault(name, []).append(v)
    for g, d in per.items():
        tot = sum(sum(v) for v in d.values())
        print(f"generation {g}: {tot:.3f} ms of kernel time (serialised, cold-cache)")
        for k, l in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            print(f"  {k:24s} n={len(l):5d} total={sum(l):9.3f} ms share={sum(l) / tot:6.3f} avg={sum(l) / len(l) * 1e3:8.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
