"""TEST-ONLY host compile of the engine sources (g++ -DOMC_EMU) into
tests/_emu/libomcgpu_emu.so.

This exists so that the CPU test tier (-m "not gpu") can check the host-side
scheduler, the work-list logic and the device functions against the oracle
without a GPU.  It is never loaded by the openmc_b200 package and is not a
fallback: the product library fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "openmc_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "_emu")
LIB = os.path.join(OUT_DIR, "libomcgpu_emu.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in ("engine.cu", "device.h", "common.h", "comm.h", "cuda_emu.h")]
    deps.append(os.path.join(ROOT, "include", "omcgpu.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) < os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DOMC_EMU", "-DOMC_EXTENDED=1", "-ffp-contract=off",
           "-fno-fast-math", "-x", "c++", os.path.join(CSRC, "engine.cu"), "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
