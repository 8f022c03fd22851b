"""Make the five edits of INTEGRATION.md to COPIES of the reference's src/simulation.cpp and
src/finalize.cpp (the reference tree itself is read-only and is never modified; the copies go to
a build directory that is not under version control).

    python apply_hooks.py /root/reference/src <out_dir>

Each edit is anchored on ONE line of the reference; the script fails loudly if an anchor is
missing or ambiguous, so a changed upstream cannot be patched silently wrong.
"""
import os
import sys

INCLUDE = '#include "omcgpu_host.h" // B200 transport path (openmc_b200/host)\n'


def edit(text, anchor, nth, replacement, path):
    lines = text.split("\n")
    hits = [i for i, l in enumerate(lines) if l.strip() == anchor]
    if len(hits) <= nth:
        raise SystemExit(f"{path}: anchor {anchor!r} (occurrence {nth}) not found")
    i = hits[nth]
    indent = lines[i][:len(lines[i]) - len(lines[i].lstrip())]
    lines[i] = "\n".join(indent + r for r in replacement)
    return "\n".join(lines)


def main(src, out):
    os.makedirs(out, exist_ok=True)
    p = os.path.join(src, "simulation.cpp")
    s = open(p).read()
    if s.count("if (settings::event_based) {") != 2:
        raise SystemExit(p + ": expected exactly two `if (settings::event_based) {`")
    # openmc_next_batch (second occurrence first, so the first keeps its index)
    s = edit(s, "if (settings::event_based) {", 1,
             ["if (gpu::enabled()) {", "  gpu::transport_generation();", "} else if (settings::event_based) {"], p)
    # openmc_simulation_init
    s = edit(s, "if (settings::event_based) {", 0,
             ["if (gpu::enabled()) {", "  gpu::simulation_init();", "} else if (settings::event_based) {"], p)
    # finalize_generation: the device has sorted and combed the bank
    s = edit(s, "sort_fission_bank();", 0, ["if (!gpu::bank_synchronized())", "  sort_fission_bank();"], p)
    s = edit(s, "synchronize_bank();", 0, ["if (!gpu::bank_synchronized())", "  synchronize_bank();"], p)
    # ... and the fission bank stayed there: its Shannon entropy comes from the device's site counts
    s = edit(s, "shannon_entropy();", 0,
             ["{", "  if (gpu::bank_synchronized())", "    gpu::shannon_entropy();", "  else", "    shannon_entropy();", "}"], p)
    s = edit(s, '#include "openmc/simulation.h"', 0, ['#include "openmc/simulation.h"', INCLUDE.rstrip("\n")], p)
    open(os.path.join(out, "simulation.cpp"), "w").write(s)

    p = os.path.join(src, "finalize.cpp")
    s = open(p).read()
    s = edit(s, "if (settings::event_based) {", 0, ["gpu::simulation_finalize();", "if (settings::event_based) {"], p)
    s = edit(s, '#include "openmc/finalize.h"', 0, ['#include "openmc/finalize.h"', INCLUDE.rstrip("\n")], p)
    open(os.path.join(out, "finalize.cpp"), "w").write(s)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
