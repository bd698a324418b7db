"""Builds tests/emu/_build/libteal_emu.so: the production kernel sources compiled as C++ against the CPU emulation
shim (cuda_emu.h).  Test infrastructure only -- g++, no nvcc, no GPU."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "teal_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libteal_emu.so")
KERNEL_SOURCES = ["api.cu", "gemv.cu", "decode_kernels.cu"]
CUDA_INC = os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-DTEAL_EMU", "-pthread", "-I", CUDA_INC, "-I", HERE,
         "-include", os.path.join(HERE, "cuda_emu.h"), "-Wno-attributes", "-Wno-unknown-pragmas"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cpp"))]
    d.append(os.path.join(ROOT, "include", "teal_b200.h"))
    return d


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in _deps()):
        return LIB
    objs = []
    jobs = []
    for src in [os.path.join(CSRC, f) for f in KERNEL_SOURCES] + [os.path.join(HERE, "emu_lib.cpp"), os.path.join(HERE, "emu_runtime.cpp")]:
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        objs.append(obj)
        jobs.append((src, subprocess.Popen(["g++", *FLAGS, "-x", "c++", "-c", src, "-o", obj], stdout=subprocess.PIPE,
                                           stderr=subprocess.STDOUT, text=True)))
    for src, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n{out[-6000:]}")
    res = subprocess.run(["g++", "-shared", "-pthread", "-o", LIB, *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
