#!/usr/bin/env python
"""Benchmark of the event-based transport path (BASELINE.json metric:
active-batch particles/s; macro-xs lookups/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle on the host cores

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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from openmc_b200 import lib, source
    from openmc_b200.csrc import build as build_ext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from openmc_b200.comm import TorchComm
        comm = TorchComm(on_gpu=True)
    build_ext.build()
    model, box, of, workload = build_workload(args)
    eng = lib.Engine()
    n_local = args.particles
    n_global = n_local * world
    eng.init(model, local_rank, n_local)
    work_lo = rank * n_local
    src = source.sample_initial_source(eng, n_local, box, of, id_offset=work_lo)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident loop ("value"): source bank stays in HBM between generations
    sim = lib.Simulation(eng, model, n_global, args.warmup + args.steps, args.warmup, src,
                         device=local_rank, resident=True, comm=comm)
    sim.engine.init = lambda *a, **k: None          # engine already initialised above
    sim.simulation_init()
    eng.set_source_bank(src)
    for _ in range(args.warmup):
        sim.next_batch()
    st0 = dict(sim.stats)
    l0 = eng.kernel_launches
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.next_batch()
    barrier()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    launches = eng.kernel_launches - l0
    st = {k: sim.stats[k] - st0[k] for k in st0}
    if world > 1:
        t = torch.tensor([dt, st["ms_transport"], st["ms_xs"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, st["ms_transport"], st["ms_xs"] = (float(x) for x in t)
        c = torch.tensor([st["n_xs_lookups"], st["n_events"], launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        st["n_xs_lookups"], st["n_events"], launches = (int(x) for x in c)
    value = n_global * args.steps / dt
    k_res = sim.keff_combined()

    # ---- end to end through host buffers ("e2e"): the reference-facing call
    e2e = None
    if not args.no_e2e:
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
        for s in range(args.steps):
            gi = eng.gen_in(n_local, work_lo, (gen0 + s) * n_global, keff, True)
            h2d += bank.nbytes
            if world == 1:
                out = eng.next_generation(gi, bank, gen0 + s + 1)
                k_gen = out.global_tally[2] / n_local
            else:
                eng.set_source_bank(bank)
                out = eng.run_generation(gi)
                g = comm.allreduce_sum(np.array(out.global_tally[:] + [out.total_weight]))
                comm.synchronize_bank(sim, gen0 + s + 1, out.n_fission)
                eng.get_source_bank(n_local, out=bank)
                k_gen = g[2] / n_global
            d2h += bank.nbytes + 96
            keff = k_gen
        barrier()
        dt_e = time.perf_counter() - t0
        eng.host_unregister(bank)
        if world > 1:
            t = torch.tensor([dt_e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_e = float(t[0])
        e2e = {"value": n_global * args.steps / dt_e, "unit": "particles/s",
               "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
               "note": ("omcgpu_next_generation on the page-locked host SourceSite bank: H2D source, event "
                        "loop, device sort + comb, D2H next source" if world == 1 else
                        "per rank: H2D of the page-locked host source bank, event loop, device comb + NCCL "
                        "exchange of the sites that change owner, D2H of the next source bank")}

    # ---- roofline of the dominant kernel (k_xs) from live CUDA-event timings
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_nuc = np.diff(np.asarray(model["mat_nuc_ofs"]))
    roof = None
    if bool(model["run_CE"]) and st["ms_xs"] > 0:
        # algorithmic bytes per item: 60 + 88 * n_nuclides(material); the
        # per-material item counts come from the lookup micro-benchmark below
        roof = xs_roofline(eng, model, min(n_local, 1_000_000), peak, peaks, box, args.workload)
        roof["in_loop"] = {"what": "device time of the event kernels that contain the lookup (k_xs_advance; "
                                   "k_surface_xa + k_collide_xa in sweep mode 2), all generations timed",
                           "ms_total": st["ms_xs"],
                           "share_of_transport": st["ms_xs"] / max(st["ms_transport"], 1e-9),
                           "lookups": st["n_xs_lookups"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu = cpu_baseline(args, src, args.cpu_sample, cores)

    if rank == 0:
        line = {
            "metric": "particles/sec (active)", "value": value, "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "particles_per_generation_per_gpu": n_local,
                       "global_particles_per_generation": n_global,
                       "l2": "working set (cross-section tables + particle state + flight records, > 0.4 GB "
                             "at 1e6 histories) exceeds L2; no flush",
                       "tallies": "the config's tallies, active (C2: 100x100 mesh {flux,fission,nu-fission} "
                                  "+ 500-bin spectrum)",
                       "parallelism": f"particles split over {world} GPU(s), replicated data"},
            "device_ms_per_step": st["ms_transport"] / args.steps,
            "events_per_particle": st["n_events"] / (n_global * args.steps),
            "xs_lookups_per_s": st["n_xs_lookups"] / (st["ms_xs"] * 1e-3) if st["ms_xs"] > 0 else None,
            "k_combined": [k_res[0], k_res[1]], "n_lost": st["n_lost"],
            "gpu_launches": launches, "clocks": clk, "e2e": e2e, "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    eng.finalize()
    if world > 1:
        dist.destroy_process_group()


def xs_roofline(eng, model, n, peak, peaks, box, args_workload="pincell_ce"):
    """macro-xs lookup kernel alone on a device-resident work list (E log-uniform over
    the library's range, materials by their volume fractions inside the source box --
    found with the device find_cell probe on uniform points): achieved algorithmic GB/s."""
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
    alg_bytes = float(np.sum(60 + 88 * n_nuc[mat]))
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
    # DRAM bytes of one such launch from the committed `ncu --set full` capture of this very
    # command (profiles/traffic.json, written from profiles/*_probe_xs_ncu.txt); None if absent
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(args_workload)
        if ent and int(ent.get("lookups", 0)) == int(n):
            traffic = float(ent["dram_bytes_per_launch"])
    except Exception:
        pass
    return {"bound": "hbm", "kernel": "k_probe_xs (macro-xs lookup, same device function as k_xs)",
            "achieved": achieved, "peak": peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": t * 1e3,
            "lookups_per_s": n / t, "nuclide_lookups_per_s": float(np.sum(n_nuc[mat])) / t}


def reference_arm(args, rank, world):
    """--impl reference: the CPU restatement (oracle) on all host cores; the
    reference itself needs HDF5 + cmake and cannot be built here."""
    if rank != 0:
        return
    from oracle.binding import Oracle
    model, box, of, workload = build_workload(args)
    o = Oracle(model)
    cores = os.cpu_count() or 1
    n = args.cpu_sample
    src = o.sample_initial_source(n, box, of)
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_baseline(args, src, n, cores)
        if s >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    line = {"impl": "reference", "metric": "particles/sec (active)", "value": v,
            "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "particles_per_step": n},
            "cpu_baseline": {"value": v, "unit": "particles/s", "cores": vals[0]["cores"],
                             "kind": "port", "sample": vals[0]["sample"]},
            "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
