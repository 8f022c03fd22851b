#!/usr/bin/env python
"""Benchmark of the event-based transport path (BASELINE.json metric:
active-batch particles/s; macro-xs lookups/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own binary on the host cores

A "step" is ONE active generation of the workload: initialise n histories from
the source bank, run the event loop to completion with the tallies active, sort
the fission bank and comb it into the next source bank (and, for N > 1, exchange
it between ranks).  Workload at N=1: config C2 of BASELINE.json -- the UO2
pincell, continuous energy, synthetic 10-nuclide x 100k-point library, with the
100x100 mesh {flux, fission, nu-fission} tally and the 500-bin flux spectrum.
Weak scaling: every GPU owns --particles histories per generation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--particles", type=int, default=4_000_000,
                    help="histories per generation per GPU (throughput = 17.5 ms fixed + 99 ms per 1e6 on C2: "
                         "profiles/r16_matrix.log; the survey's range is 1e6-1e7)")
    ap.add_argument("--n-grid", type=int, default=100_000, help="energy-grid points per nuclide")
    ap.add_argument("--workload", default="pincell_ce",
                    choices=["pincell_ce", "pincell_mg", "assembly_ce", "c5g7_core", "pwr_core_ce"],
                    help="pincell_ce = config C2 (the bench line); the others are C1, C3, C4, C5")
    ap.add_argument("--n-fuel-nuclides", type=int, default=300, help="pwr_core_ce: nuclides in the fuel")
    ap.add_argument("--cpu-sample", type=int, default=160_000, help="histories of the cpu_baseline sample (~25 s of CPU work)")
    ap.add_argument("--ref-particles", type=int, default=0,
                    help="histories per batch of the reference binary's runs (default 10000 per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the C5 / C1 / C3 / C4 sub-lines (N = 1 runs of the default workload add them)")
    return ap.parse_args()


def build_workload_description(args):
    w = args.workload
    name = {"pincell_ce": f"UO2 pincell CE, synthetic library 10 nuclides x {args.n_grid} pts (config C2)",
            "assembly_ce": f"17x17 assembly CE, synthetic library x {args.n_grid} pts, pin-wise mesh tally (config C3)",
            "pwr_core_ce": f"Hoogenboom-Martin full core CE, {args.n_fuel_nuclides}-nuclide fuel x {args.n_grid} pts (config C5)",
            "c5g7_core": "C5G7-style 2-D core MG, 510x510 mesh tally (config C4)",
            "pincell_mg": "pincell_multigroup 7-group C5G7 (config C1)"}[w]
    return None, None, None, name


def build_workload(args):
    from openmc_b200 import examples, mgxs, nuclide
    w = args.workload
    if w == "pincell_ce":
        m, libx, box, of = examples.pincell_ce(n_grid=args.n_grid)
        return nuclide.build_ce_model(m, libx), box, of, \
            f"UO2 pincell CE, synthetic library 10 nuclides x {args.n_grid} pts (config C2)"
    if w == "assembly_ce":
        m, libx, box, of = examples.assembly_ce(n_grid=args.n_grid)
        return nuclide.build_ce_model(m, libx), box, of, \
            f"17x17 assembly CE, synthetic library x {args.n_grid} pts, pin-wise mesh tally (config C3)"
    if w == "pwr_core_ce":
        m, libx, box, of = examples.pwr_core_ce(n_grid=args.n_grid, n_fuel_nuclides=args.n_fuel_nuclides)
        return nuclide.build_ce_model(m, libx), box, of, \
            f"Hoogenboom-Martin full core CE, {args.n_fuel_nuclides}-nuclide fuel x {args.n_grid} pts (config C5)"
    if w == "c5g7_core":
        m, libx, box, of = examples.c5g7_core()
        return mgxs.build_mg_model(m, libx), box, of, "C5G7-style 2-D core MG, 510x510 mesh tally (config C4)"
    m, libx, box, of = examples.pincell_multigroup()
    return mgxs.build_mg_model(m, libx), box, of, "pincell_multigroup 7-group C5G7 (config C1)"


# ------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples if len(s) >= 6
                          for k in range(4) if s[2 + k].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------
# CPU baseline: the oracle (port) on the host cores, bounded sample
# ------------------------------------------------------------------------------
def _cpu_worker(job):
    model_args, src, id_offset, seed_base = job
    from oracle.binding import Oracle
    from openmc_b200 import lib
    args = argparse.Namespace(**model_args)
    model, _, _, _ = build_workload(args)
    o = Oracle(model)
    gi = lib.Engine.gen_in(None, src.size, id_offset, seed_base, 1.0, True)
    t0 = time.perf_counter()
    _, out = o.transport_generation(gi, src)
    return time.perf_counter() - t0, int(out.n_xs_lookups), int(out.n_events)


def cpu_baseline(args, src, n_sample, cores):
    """time the oracle on `cores` host processes, each transporting a slice of the sample"""
    import multiprocessing as mp
    n_sample = min(n_sample, src.size)
    per = max(n_sample // cores, 1)
    margs = dict(workload=args.workload, n_grid=args.n_grid, n_fuel_nuclides=args.n_fuel_nuclides)
    jobs = [(margs, src[k * per:(k + 1) * per].copy(), k * per, 0) for k in range(cores)]
    jobs = [j for j in jobs if j[1].size]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)     # workers run concurrently; model build excluded
    n = sum(j[1].size for j in jobs)
    return {"value": n / wall, "unit": "particles/s", "cores": len(jobs), "kind": "port",
            "sample": f"{n} histories of one generation of the same workload (tallies active), "
                      f"{len(jobs)} processes x oracle/oracle.cpp; total wall incl. spawn {time.perf_counter() - t0:.1f}s",
            "xs_lookups_per_s": sum(r[1] for r in res) / wall}


# ------------------------------------------------------------------------------
# byte model of SURVEY.md section 8(d) (the figure `roofline.achieved` is built on)
# ------------------------------------------------------------------------------
def bytes_xs(n_nuclides_total, n_lookups):
    """B_macro(n) = 60 + 116 n per macro lookup actually evaluated"""
    return 60.0 * n_lookups + 116.0 * n_nuclides_total


def state_bytes(levels):
    """per-event particle-state traffic (SoA, `levels` coordinate levels)"""
    L1 = max(levels, 1) - 1
    return {"advance": (200 + 72 * L1) + (150 + 48 * L1), "cross_surface": 160 + 120, "collide": 140 + 120}


RED_PEAK = 210e9   # measured L2 RED.F64 rate on this GPU (profiles/r13_red_peak.txt), adds per second


def kernel_rooflines(model, steps_stats, count_stats, peak, fused=True):
    """Per in-loop kernel: algorithmic bytes (section 8(d) model) of ALL its launches in the timed
    steps / the sum of their durations (CUDA events on the kernel's own stream).  The lookup
    statistics (share of events whose lookup is evaluated, nuclides per lookup) and the RED count
    come from one extra, untimed generation run with the engine's counters on."""
    L = int(model["n_coord_levels"])
    sb = state_bytes(L)
    st, cs = steps_stats, count_stats
    out = {}

    def entry(name, items, ms, launches, bytes_total, extra=None):
        if ms <= 0 or items <= 0:
            return
        e = {"items": int(items), "launches": int(launches), "ms_total": ms,
             "avg_launch_ms": ms / max(launches, 1),
             "algorithmic_bytes_per_item": bytes_total / items,
             "achieved": bytes_total / (ms * 1e-3) / 1e9, "unit": "GB/s", "peak": peak}
        e["frac"] = e["achieved"] / peak
        if e["frac"] > 1.0:
            e["note"] = ("algorithmic bytes count 116 B per (lookup, nuclide) as if every lookup fetched its own sectors; "
                         "on the (material, energy)-sorted list neighbouring lanes share hash windows, grid points and "
                         "xs rows, so the DRAM traffic is several times lower (ncu: profiles/r40_c5_ncu.txt, 701 MB in "
                         "471 us = 1.5 TB/s for this kernel): a fraction above 1 means the sort removed the traffic, "
                         "not that HBM was exceeded")
        e["items_per_s"] = items / (ms * 1e-3)
        if extra:
            e.update(extra)
        out[name] = e

    # crossing kernel: cross_surface + [lookup] + advance per item (dead particles make no flight;
    # they are a few per cent and are charged the full event)
    sfx = "_xa" if fused else ""
    for k, (name, ev, items, ms, nl) in enumerate((
            ("k_surface" + sfx, "cross_surface", st["n_items_surface"], st["ms_surface"], st["n_launch_surface"]),
            ("k_collide" + sfx, "collide", st["n_items_collide"], st["ms_collide"], st["n_launch_collide"]))):
        c_items = max(cs["n_items_surface"] if k == 0 else cs["n_items_collide"], 1)
        if fused:
            lookups_per_item = cs["lookups_done"][k] / c_items
            nuc_per_item = cs["nuclide_lookups"][k] / c_items
            b = items * (sb[ev] + sb["advance"] + 120.0) + bytes_xs(nuc_per_item * items, lookups_per_item * items)
            extra = {"lookups_per_item": lookups_per_item,
                     "nuclides_per_lookup": cs["nuclide_lookups"][k] / max(cs["lookups_done"][k], 1),
                     "model": f"{ev} {sb[ev]} B + advance {sb['advance']} B + flight record 120 B + "
                              "(60 + 116 n) B per lookup evaluated"}
        else:       # merged sweep: the event alone (lookup and flight are kernels of their own)
            b = items * float(sb[ev])
            extra = {"model": f"{ev} {sb[ev]} B"}
        if k == 1:   # banked fission sites: 104 B each
            b += 104.0 * st["n_fission"]
        entry(name, items, ms, nl, b, extra)
    flights = st["n_items_surface"] + st["n_items_collide"] + st["n_first_flights"]
    c_flights = max(cs["n_items_surface"] + cs["n_items_collide"] + cs["n_first_flights"], 1)
    red_per_flight = cs["n_red"] / c_flights
    if st["ms_score"] > 0:
        reds = red_per_flight * flights
        entry("k_score", flights, st["ms_score"], st["n_launch_score"], flights * 120.0 + 8.0 * reds,
              {"red_per_item": red_per_flight, "red_per_s": reds / (st["ms_score"] * 1e-3),
               "red_frac_of_measured_peak": reds / (st["ms_score"] * 1e-3) / RED_PEAK,
               "model": "flight record 120 B + 8 B per RED.F64 (bin, score); bound is the L2 atomic rate "
                        f"({RED_PEAK / 1e9:.0f} G/s measured, profiles/r13_red_peak.txt), not HBM"})
    if st.get("ms_xs_kernel", 0) > 0 and cs["lookups_done"][2] > 0:      # stand-alone lookup kernel (merged sweep)
        items = st["n_xs_kernel_items"]
        per_item_l = cs["lookups_done"][2] / max(cs["n_xs_kernel_items"], 1)
        per_item_n = cs["nuclide_lookups"][2] / max(cs["n_xs_kernel_items"], 1)
        entry("k_xs", items, st["ms_xs_kernel"], st["n_launch_xs"], bytes_xs(per_item_n * items, per_item_l * items),
              {"lookups_per_item": per_item_l, "nuclide_lookups_per_s": per_item_n * items / (st["ms_xs_kernel"] * 1e-3),
               "nuclides_per_lookup": cs["nuclide_lookups"][2] / max(cs["lookups_done"][2], 1),
               "model": "(60 + 116 n) B per lookup evaluated"})
    return out


_STAT_KEYS = ("ms_transport", "ms_xs", "ms_surface", "ms_collide", "ms_score", "n_launch_surface",
              "n_launch_collide", "n_launch_score", "n_items_surface", "n_items_collide", "n_events",
              "n_xs_lookups", "n_collisions", "n_crossings", "n_lost", "n_fission", "n_red")


def _accumulate(acc, out, n_first):
    for k in _STAT_KEYS:
        acc[k] = acc.get(k, 0) + getattr(out, k)
    acc["n_first_flights"] = acc.get("n_first_flights", 0) + n_first
    for name in ("lookups_done", "nuclide_lookups"):
        v = acc.setdefault(name, [0, 0, 0])
        for k in range(3):
            v[k] += getattr(out, name)[k]
    return acc


def load_traffic(workload_key):
    """DRAM bytes per list item of each in-loop kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, written by profiles/ncu_traffic.py from the .ncu-rep of this command)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload_key)
    except Exception:
        return None


def measure(args, workload_key, n_local, steps, warmup, rank, local_rank, world, comm, with_e2e, peak, peaks):
    """one workload on this rank's GPU: resident loop (`value`), host-buffer loop (`e2e`), per-kernel rooflines"""
    import torch
    from openmc_b200 import lib, source

    wargs = argparse.Namespace(**vars(args))
    wargs.workload = workload_key
    model, box, of, workload = build_workload(wargs)
    eng = lib.Engine()
    n_global = n_local * world
    eng.init(model, local_rank, n_local)
    work_lo = rank * n_local
    src = source.sample_initial_source(eng, n_local, box, of, id_offset=work_lo)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            comm.barrier()          # the library's own all-gather over NCCL
            torch.cuda.synchronize()

    # ---- resident loop ("value"): source bank stays in HBM between generations
    sim = lib.Simulation(eng, model, n_global, warmup + steps + 1, warmup, src,
                         device=local_rank, resident=True, comm=comm)
    sim.engine.init = lambda *a, **k: None          # engine already initialised above
    sim.simulation_init()
    eng.set_source_bank(src)
    acc = {}
    sim.on_generation = lambda out: _accumulate(acc, out, n_local)
    for _ in range(warmup):
        sim.next_batch()
    acc.clear()
    st0 = dict(sim.stats)
    l0 = eng.kernel_launches
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.next_batch()
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    launches = eng.kernel_launches - l0
    st = dict(acc)
    st["ms_xs_kernel"] = st["ms_xs"] if eng.sweep_kind == 0 else 0.0
    # one more, untimed generation with the engine's counters on (lookup statistics, RED count)
    acc.clear()
    eng.set_option("count_lookups", 1)
    sim.next_batch()
    eng.set_option("count_lookups", 0)
    cs = dict(acc)
    sim.on_generation = None
    if eng.sweep_kind == 0:    # merged sweep: every alive history is an item of the lookup kernel
        st["n_xs_kernel_items"] = st["n_xs_lookups"]; st["n_launch_xs"] = st["n_launch_surface"]
        cs["n_xs_kernel_items"] = cs["n_xs_lookups"]
    if world > 1:
        dt = comm.max_over_ranks(dt)
        st["ms_transport"] = comm.max_over_ranks(st["ms_transport"] * 1e-3) * 1e3
        n_events_all = comm.sum_over_ranks(st["n_events"])
        launches = comm.sum_over_ranks(launches)
    else:
        n_events_all = st["n_events"]
    value = n_global * steps / dt
    k_res = sim.keff_combined()

    # ---- end to end through host buffers ("e2e"): the reference-facing call
    e2e = None
    if with_e2e:
        # The reference-facing calls on HOST buffers: libopenmc's simulation::source_bank
        # (page-locked once, as allocate_banks would) goes H2D, the generation runs, and
        # the NEXT source bank comes back D2H.  N = 1: omcgpu_next_generation (device sort
        # + comb = transport_event_based + synchronize_bank).  N > 1: omcgpu_set_source_bank,
        # omcgpu_run_generation, the inter-rank comb + exchange of synchronize_bank on the
        # device banks (comm.py), omcgpu_get_source_bank into the same host bank.
        bank = np.ascontiguousarray(eng.get_source_bank(n_local))
        eng.host_register(bank)
        gen0 = sim.current_batch
        h2d = d2h = 0
        barrier()
        t0 = time.perf_counter()
        keff = sim.keff
        for s in range(steps):
            gi = eng.gen_in(n_local, work_lo, (gen0 + s) * n_global, keff, True)
            h2d += bank.nbytes
            if world == 1:
                out = eng.next_generation(gi, bank, gen0 + s + 1)
                k_gen = out.global_tally[2] / n_local
            else:
                eng.set_source_bank(bank)
                out = eng.run_generation(gi)
                # omcgpu_sync_bank: comb + NCCL exchange + sum of the globals + the D2H mirror of
                # the new source into the host bank, one library call
                g = comm.sync_bank(sim, gen0 + s + 1, np.array(out.global_tally[:] + [out.total_weight]),
                                   host_bank=bank)
                k_gen = g[2] / n_global
            d2h += bank.nbytes + 96
            keff = k_gen
        barrier()
        dt_e = time.perf_counter() - t0
        eng.host_unregister(bank)
        if world > 1:
            dt_e = comm.max_over_ranks(dt_e)
        e2e = {"value": n_global * steps / dt_e, "unit": "particles/s",
               "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps,
               "note": ("omcgpu_next_generation on the page-locked host SourceSite bank: H2D source, event "
                        "loop, device sort + comb, D2H next source" if world == 1 else
                        "per rank: H2D of the page-locked host source bank (omcgpu_set_source_bank), event loop "
                        "(omcgpu_run_generation), omcgpu_sync_bank = device comb + NCCL exchange of the column "
                        "ranges that change owner + NCCL sum of the globals + D2H of the next source bank")}

    # ---- rooflines of the in-loop kernels from the live CUDA-event timings
    kernels = kernel_rooflines(model, st, cs, peak, eng.fused_lookup)
    traffic = load_traffic(workload_key) or {}
    for name, e in kernels.items():
        t = traffic.get(name)
        e["traffic"] = (t["dram_bytes_per_item"] * e["items"] / e["launches"]) if t else None
        if t:
            e["ncu"] = {k: t[k] for k in t if k != "dram_bytes_per_item"}
            e["traffic_over_algorithmic"] = t["dram_bytes_per_item"] / e["algorithmic_bytes_per_item"]
    roof = None
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["ms_total"])
        # the headline object describes the event kernel that takes the most device time; k_score is
        # bound by L2 atomics (see its red_frac_of_measured_peak), the event kernels by HBM gathers
        ev_k = [k for k in kernels if k != "k_score"]
        head = max(ev_k, key=lambda k: kernels[k]["ms_total"]) if ev_k else dom
        e = kernels[head]
        roof = {"bound": "hbm", "kernel": head, "achieved": e["achieved"],
                "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": e["frac"], "traffic": e["traffic"],
                "algorithmic_bytes_per_launch": e["algorithmic_bytes_per_item"] * e["items"] / e["launches"],
                "avg_launch_ms": e["avg_launch_ms"],
                "how": "algorithmic bytes (SURVEY 8(d) model: per-event state bytes + (60 + 116 n) per lookup "
                       "evaluated) of all launches of the kernel in the timed steps / the sum of their "
                       "durations, CUDA events on the kernel's own stream",
                "largest_kernel_by_time": dom, "kernels": kernels,
                "step_check_GBps": sum(k["achieved"] * k["ms_total"] for k in kernels.values()) /
                max(st["ms_transport"], 1e-9)}
        if bool(model["run_CE"]) and workload_key == args.workload and world == 1:
            roof["lookup_probe_L2_resident" if workload_key == "pincell_ce" else "lookup_probe"] = \
                xs_probe(eng, model, min(n_local, 1_000_000), peak, box)
    res = {"workload": workload, "value": value, "ms_per_step": 1e3 * dt / steps,
           "device_ms_per_step": st["ms_transport"] / steps, "n_local": n_local, "n_global": n_global,
           "events_per_particle": n_events_all / (n_global * steps), "k_combined": [k_res[0], k_res[1]],
           "n_lost": st["n_lost"], "gpu_launches": launches, "clocks": clk, "e2e": e2e, "roofline": roof,
           "sweep_mode": {2: "paired lists, lookup + flight fused into the crossing / collision kernels",
                          1: "merged alive list, one lookup + flight kernel",
                          0: "merged alive list, stand-alone lookup kernel over the (material, energy)-sorted list"
                          }[eng.sweep_kind],
           "src": src}
    eng.finalize()
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    from openmc_b200.csrc import build as build_ext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        # one process per GPU; every inter-rank step is the library's own (NCCL behind the C ABI,
        # csrc/comm.h).  torch only lends its rendezvous key-value store (plain TCP, no process
        # group) to hand the NCCL unique id from rank 0 to the others.
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        pin_to_gpu_numa_node(local_rank)
        from torch.distributed import rendezvous
        from openmc_b200.comm import NativeComm
        store, _, _ = next(rendezvous("env://", rank, world))
        comm = NativeComm.over_nccl(rank, world, store)
    if rank == 0:
        build_ext.build()
    if world > 1:               # the other ranks load the library only once rank 0 has (re)built it
        if rank == 0:
            store.set("omcgpu_built", b"1")
        else:
            store.get("omcgpu_built")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    r = measure(args, args.workload, args.particles, args.steps, args.warmup, rank, local_rank, world, comm,
                not args.no_e2e, peak, peaks)

    # ---- strong-scaling point (N > 1): the SAME global generation as the N = 1 line (--particles
    # histories in total, --particles / N per GPU) through the same NCCL path
    strong = None
    if world > 1 and not args.no_other_configs:
        try:
            o = measure(args, args.workload, max(args.particles // world, 1), 3, 3, rank, local_rank, world, comm,
                        False, peak, peaks)
            strong = {"value": o["value"], "unit": "particles/s", "ms_per_step": o["ms_per_step"],
                      "device_ms_per_step": o["device_ms_per_step"], "global_particles_per_generation": o["n_global"],
                      "particles_per_generation_per_gpu": o["n_local"], "steps": 3, "warmup": 3,
                      "note": "fixed total work: the N = 1 generation split over N GPUs"}
        except Exception as exc:
            strong = {"error": repr(exc)[:300]}

    # ---- the other configurations of BASELINE.json (N = 1 only): C5 is the only one whose tables
    # (1.8 GB) do not fit L2, i.e. the HBM-gather case of north_star
    others = {}
    if world == 1 and not args.no_other_configs and args.workload == "pincell_ce":
        for key, tag, n_part in (("pwr_core_ce", "C5", 2_000_000), ("pincell_mg", "C1", 2_000_000),
                                 ("assembly_ce", "C3", 2_000_000), ("c5g7_core", "C4", 2_000_000)):
            try:
                o = measure(args, key, n_part, 3, 3, rank, local_rank, world, None, False, peak, peaks)
                o.pop("src")
                others[tag] = {k: o[k] for k in ("workload", "value", "ms_per_step", "device_ms_per_step", "n_local",
                                                 "events_per_particle", "k_combined", "n_lost", "sweep_mode",
                                                 "roofline")}
                others[tag]["unit"] = "particles/s"
                others[tag]["steps"], others[tag]["warmup"] = 3, 3
            except Exception as exc:     # a sub-line must not take the bench line down
                others[tag] = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if reference_binary_available():
            cpu = reference_binary_baseline(args, args.ref_particles or 10000 * cores, 2, 1, cores)
        else:
            cpu = cpu_baseline(args, r["src"], args.cpu_sample, cores)

    dropin = None
    if rank == 0 and world == 1 and not args.no_e2e and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "openmc_gpu")):
        try:
            dropin = dropin_rate(args, args.particles, 3, 2, os.cpu_count() or 1)
        except Exception as exc:
            dropin = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {
            "metric": "particles/sec (active)", "value": r["value"], "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["workload"], "particles_per_generation_per_gpu": r["n_local"],
                       "global_particles_per_generation": r["n_global"],
                       "l2": "working set (cross-section tables + particle state + flight records, > 0.4 GB "
                             "at 1e6 histories) exceeds L2; no flush",
                       "tallies": "the config's tallies, active (C2: 100x100 mesh {flux,fission,nu-fission} "
                                  "+ 500-bin spectrum)",
                       "parallelism": f"particles split over {world} GPU(s), replicated data",
                       "sweep_mode": r["sweep_mode"]},
            "device_ms_per_step": r["device_ms_per_step"],
            "events_per_particle": r["events_per_particle"],
            "k_combined": r["k_combined"], "n_lost": r["n_lost"],
            "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "e2e": r["e2e"], "roofline": r["roofline"],
            "cpu_baseline": cpu, "other_configs": others or None, "dropin": dropin, "strong_scaling": strong,
        }
        print(json.dumps(line))


def pin_to_gpu_numa_node(index):
    """best effort: run this process (and first-touch its page-locked banks) on the CPU cores of
    the NUMA node the GPU hangs off -- the host-buffer leg moves 0.8 GB per generation per GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def xs_probe(eng, model, n, peak, box):
    """macro-xs lookup kernel ALONE on a device-resident work list (E log-uniform over the
    library's range, materials by their volume fractions inside the source box -- found with
    the device find_cell probe on uniform points).  A secondary entry: on C2 its 48 MB of tables
    sit in L2 (DRAM traffic far below the algorithmic bytes), so this is L2 gather throughput."""
    import ctypes as C
    import torch
    g = np.random.default_rng(7)
    e_min, e_max = float(model["energy_min"]), float(model["energy_max"])
    E = np.exp(g.uniform(np.log(e_min), np.log(e_max), n))
    box = np.asarray(box, float)
    xyz = box[:3] + g.uniform(size=(n, 3)) * (box[3:] - box[:3])
    mat = eng.find_cell(xyz)[2].astype(np.int32)
    good = mat[mat >= 0]
    mat = np.where(mat >= 0, mat, good[0] if good.size else 0).astype(np.int32)
    n_nuc = np.diff(np.asarray(model["mat_nuc_ofs"]))
    alg_bytes = float(np.sum(60 + 116 * n_nuc[mat]))
    dE = torch.from_numpy(E).cuda()
    dm = torch.from_numpy(mat).cuda()
    out = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    ms = C.c_float(0)
    best = []
    for it in range(8):
        eng._check(eng.lib.omcgpu_calculate_xs_device(
            C.c_void_p(dE.data_ptr()), C.c_void_p(dm.data_ptr()), C.c_int64(n),
            C.c_void_p(out.data_ptr()), C.byref(ms)))
        if it >= 3:
            best.append(ms.value)
    t = float(np.mean(best)) * 1e-3
    achieved = alg_bytes / t / 1e9
    return {"kernel": "k_probe_xs (macro-xs lookup alone, unsorted random energies)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": t * 1e3,
            "lookups_per_s": n / t, "nuclide_lookups_per_s": float(np.sum(n_nuc[mat])) / t,
            "model": "(60 + 116 n) B per lookup"}


def reference_binary_available():
    from oracle import refrun
    return refrun.available()


def dropin_rate(args, n_per_batch, active, inactive, threads):
    """the user-visible drop-in: the reference's own binary with the three hooks compiled in
    (oracle/_ref/openmc_gpu = unmodified libopenmc sources + openmc_b200/host), OPENMC_GPU=1, in a
    directory with model.xml + the data library; the rate is the line its own output prints.  Host
    side per generation as in libopenmc: sort_fission_bank, synchronize_bank (serial comb),
    k statistics, tally accumulation."""
    r = reference_binary_baseline(args, n_per_batch, active, inactive, threads, modes=("history",),
                                  binary="openmc_gpu", env={"OPENMC_GPU": "1", "OMCGPU_LIB": os.path.join(
                                      ROOT, "openmc_b200", "csrc", "libomcgpu.so")})
    h = r["history"]
    return {"value": h["particles_per_s"], "unit": "particles/s", "inactive_particles_per_s": h["inactive_particles_per_s"],
            "active_s": h["active_s"], "wall_s": h["wall_s"], "transport_s": h["transport_s"],
            "accumulating_tallies_s": h["accumulating_tallies_s"], "writing_statepoints_s": h["writing_statepoints_s"],
            "particles_per_generation": n_per_batch,
            "batches": [inactive, active],
            "what": "OPENMC_GPU=1 oracle/_ref/openmc_gpu (the reference's binary + hooks) on model.xml: its own "
                    "'Calculation Rate (active)'; host side (bank sort / comb, statistics, statepoint) is libopenmc's"}


def reference_binary_baseline(args, n_per_batch, active, inactive, threads, modes=("history", "event"),
                              binary="openmc", env=None):
    """time the UNMODIFIED reference (oracle/_ref/openmc: every translation unit of libopenmc compiled
    from /root/reference with OpenMP, against oracle/h5shim) on the workload: XML inputs + the synthetic
    library written by oracle/refrun.py, `inactive` + `active` batches of n_per_batch histories, all
    host threads, history-based and --event.  The rate is the reference's own "Calculation Rate
    (active)" line (src/output.cpp:457-497)."""
    import shutil
    import tempfile
    from openmc_b200 import examples
    from oracle import refrun
    w = args.workload
    d = tempfile.mkdtemp(prefix="omc_ref_")
    try:
        if w in ("pincell_ce", "assembly_ce", "pwr_core_ce"):
            if w == "pincell_ce":
                m, libx, box, of = examples.pincell_ce(n_grid=args.n_grid)
            elif w == "assembly_ce":
                m, libx, box, of = examples.assembly_ce(n_grid=args.n_grid)
            else:
                m, libx, box, of = examples.pwr_core_ce(n_grid=args.n_grid, n_fuel_nuclides=args.n_fuel_nuclides)
            refrun.write_ce_library(os.path.join(d, "data"), libx)
            xs, ce = "data/cross_sections.xml", True
        else:
            m, libx, box, of = examples.c5g7_core() if w == "c5g7_core" else examples.pincell_multigroup()
            refrun.write_mgxs_library(os.path.join(d, "mgxs.h5"), libx)
            xs, ce = "mgxs.h5", False
        with open(os.path.join(d, "model.xml"), "w") as f:
            f.write(refrun.model_xml(m, xs, n_per_batch, inactive + active, inactive, box, of, run_CE=ce,
                                     extra={"output": "<summary>false</summary><tallies>false</tallies>",
                                            # the final statepoint keeps k and the tallies, not the source bank
                                            # (416 MB at 4e6 sites: an I/O cost inside the reference's active
                                            # timer that has nothing to do with transport)
                                            "source_point": "<write>false</write>"}))
        res = {}
        for mode in modes:
            t0 = time.perf_counter()
            saved = {k: os.environ.get(k) for k in (env or {})}
            os.environ.update(env or {})
            try:
                r = refrun.run(d, threads=threads, event=(mode == "event"),
                               binary=os.path.join(ROOT, "oracle", "_ref", binary))
            finally:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
            if r["returncode"] != 0:
                raise RuntimeError("reference run failed: " + (r["stdout"][-400:] + r["stderr"][-400:]))
            res[mode] = {"particles_per_s": r["rates"]["active"], "inactive_particles_per_s": r["rates"].get("inactive"),
                         "active_s": r["times"].get("Time in active batches"), "wall_s": time.perf_counter() - t0,
                         "transport_s": r["times"].get("Time in transport only"),
                         "accumulating_tallies_s": r["times"].get("Time accumulating tallies"),
                         "writing_statepoints_s": r["times"].get("Time writing statepoints")}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    best = max(res, key=lambda k: res[k]["particles_per_s"])
    return {"value": res[best]["particles_per_s"], "unit": "particles/s", "cores": threads, "kind": "reference",
            "mode": best, "history": res.get("history"), "event": res.get("event"),
            "sample": f"the reference's own openmc binary (unmodified sources, OpenMP, CMake-Release flags -O3 -DNDEBUG), {threads} threads: "
                      f"{inactive} inactive + {active} active batches of {n_per_batch} histories of the same "
                      f"workload (XML + synthetic library), tallies active in the active batches; value = its "
                      f"'Calculation Rate (active)' in the faster of history-based and --event mode"}


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores --
    oracle/_ref/openmc (the unmodified reference built against oracle/h5shim) when it is there,
    else the oracle port."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, _, _, workload = build_workload_description(args)
    if reference_binary_available():
        n = args.ref_particles or 10000 * cores
        cpu = reference_binary_baseline(args, n, args.steps, args.warmup, cores)
        v = cpu["value"]
    else:
        from oracle.binding import Oracle
        model, box, of, workload = build_workload(args)
        o = Oracle(model)
        n = args.cpu_sample
        src = o.sample_initial_source(n, box, of)
        vals = []
        for s in range(args.warmup + args.steps):
            r = cpu_baseline(args, src, n, cores)
            if s >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        cpu = {"value": v, "unit": "particles/s", "cores": vals[0]["cores"], "kind": "port",
               "sample": vals[0]["sample"]}
    line = {"impl": "reference", "metric": "particles/sec (active)", "value": v,
            "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "particles_per_step": n},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
