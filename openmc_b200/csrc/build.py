"""Build libomcgpu.so (the product) in-tree with nvcc for sm_100a.

    python -m openmc_b200.csrc.build [--force]

-fmad=false keeps the device arithmetic free of fused multiply-adds so that it
rounds like the reference's x86-64 build (and the oracle); the path is
latency/bandwidth bound, so this costs nothing measurable.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libomcgpu.so")
# the same sources with the rarely needed features compiled in (device.h, OMC_EXTENDED: hexagonal lattices,
# secondary neutrons of integer yields): loaded instead of LIB for the models omcgpu_needs_extended_build
# names (openmc_b200/lib.py, host/omcgpu_host.cpp)
LIB_EXT = os.path.join(HERE, "libomcgpu_ext.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.h", "device.h", "comm.h", "cuda_emu.h", os.path.join("..", "..", "include", "omcgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
              "-Xptxas", "-v", "-I/usr/include", "-ldl"]


def stale():
    if not os.path.exists(LIB) or not os.path.exists(LIB_EXT):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_EXT))
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build_variant(name, defines):
    """experiment builds (openmc_b200/csrc/variants/<name>.so), selected with OMCGPU_LIB"""
    out = os.path.join(HERE, "variants", name + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = [d if d.startswith("-") else "-D" + d for d in defines]   # raw nvcc flags pass through
    cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + extra + ["-o", out] + \
          [os.path.join(HERE, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    return out


def build(force=False, verbose=False):
    if not (force or stale()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    jobs = []
    for out, extra, logname in ((LIB, [], "build.log"), (LIB_EXT, ["-DOMC_EXTENDED=1"], "build_ext.log")):
        cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + extra + ["-o", out] + \
              [os.path.join(HERE, s) for s in SOURCES]
        jobs.append((cmd, logname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for cmd, logname, proc in jobs:      # the two builds run side by side
        text, _ = proc.communicate()
        log = os.path.join(HERE, logname)
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + text)
        if proc.returncode != 0:
            sys.stderr.write(text)
            failed = log
        elif verbose:
            print(text)
    if failed:
        raise RuntimeError("nvcc failed, see " + failed)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
