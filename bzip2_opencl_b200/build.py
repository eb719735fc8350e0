"""Builds libbz2b200.so (the CUDA product, sm_100a) in-tree with nvcc.  No GPU is needed to build."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "bz2b200.cu")
LIB = os.path.join(HERE, "libbz2b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in os.listdir(d)]
    out.append(os.path.join(HERE, "..", "include", "bz2b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    build_tools(force)
    return LIB


def build_tools(force: bool = False) -> None:
    """Host programs on top of the facade headers (g++): facade_bench (bench.py's e2e_facade leg)."""
    src = os.path.join(HERE, "tools_src", "facade_bench.cpp")
    exe = os.path.join(HERE, "facade_bench")
    deps = [src, LIB] + [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))]
    if not force and os.path.exists(exe) and all(os.path.getmtime(d) <= os.path.getmtime(exe) for d in deps):
        return
    cmd = ["g++", "-std=c++17", "-O2", src, "-o", exe, "-DBZ2B200_BLOCKSIZE_DEFAULT=100000", "-I" + os.path.join(HERE, "include"),
           "-I" + os.path.join(HERE, "..", "include"), "-L" + HERE, "-lbz2b200", "-Wl,-rpath," + HERE, "-pthread"]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
