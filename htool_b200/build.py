"""Builds htool_b200/lib/libhtool_b200.so (the product: CUDA kernels + C ABI) with nvcc for sm_100a.

Usage: python -m htool_b200.build [--force]
The library is built IN-TREE so that it travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libhtool_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas",
          "-I", os.path.join(REPO, "include"), "-I", CSRC]
SOURCES = ["packer.cpp", "kernels.cu", "mkernels.cu", "generate.cu", "aca.cu", "capi.cu", "dist.cu", "gmres.cu"]
HEADERS = ["store.hpp", "packer.hpp", "kernels.cuh", "mkernels.cuh", "generate.cuh", "aca.cuh", "kernel_functions.cuh", "handle.hpp", os.path.join(REPO, "include", "htool_b200.h")]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [NVCC, *ARCH, *COMMON, "-c", s, "-o", o]
            if src.endswith(".cu"):
                cmd += ["-Xptxas", "-v"] if verbose else []
            else:
                cmd += ["-x", "cu"] if False else []
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _newer(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs, "-ldl", "-Xcompiler", "-fopenmp", "-lgomp"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
