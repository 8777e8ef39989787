"""Build libswgl_b200.so in-tree: nvcc for the CUDA device layer (sm_100a), gcc for the C host layer.

    python -m swgl_b200.build [--force]

Flags that matter for parity: -fmad=false -prec-div=true -prec-sqrt=true
-ftz=false on the device side, -ffp-contract=off on the host side (the GLSL literal parser
works in double precision and must not be contracted either).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "libswgl_b200.so")
BUILD = os.path.join(HERE, "_build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
    "-I", INC, "-I", CSRC,
]
C_FLAGS = ["-std=gnu11", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter",
           "-I", INC, "-I", CSRC]

CU_SOURCES = ["swgl_dev.cu"]
C_SOURCES = ["swgl_host.c", "swgl_glsl.c"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(INC, f) for f in os.listdir(INC)]
    objs = []
    for name in CU_SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(BUILD, name + ".o")
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.run(cmd, check=True)
        objs.append(obj)
    for name in C_SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(BUILD, name + ".o")
        if force or _newer([src] + headers, obj):
            subprocess.run([CC] + C_FLAGS + ["-c", src, "-o", obj], check=True)
        objs.append(obj)
    if force or _newer(objs, OUT):
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
        cmd = [os.environ.get("CXX", "g++"), "-shared", "-o", OUT] + objs + [
            "-Wl,-Bsymbolic", "-L" + cuda_lib, "-Wl,-rpath," + cuda_lib, "-lcudart", "-lstdc++"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
