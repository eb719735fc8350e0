#!/usr/bin/env python3
"""Build the REAL reference (Stan1slav337/Bzip2-OpenCL) for the CPU from the sources where they lie
under /root/reference, into oracle/_ref/ (git-ignored, travels to the GPU box as built .so files).

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under bzip2_opencl_b200/ may load these libraries.

Recipe:
  1. gen.cpp includes /root/reference/kernel.cpp and prints get_opencl_c_code()  -> _ref/kernel.cl
     (the OpenCL C text the reference would hand to the OpenCL JIT; one token per line).
  2. Defined-behaviour patches on that text (each pattern must occur the stated number of times):
       P1  zero-initialise tableFrequencies            (kernel.cpp:2902, uninitialised ++)
       P3  BLOCKSIZE_DEFAULT 10000 -> 100000           (only for the base-100000 build)
       P4  TRBudget passed by pointer                  (kernel.cpp:1656,2127; by-value makes lsSort dead code)
     P2 (fresh per-block freq[258], block buffer bs+1) lives in ref_shim.cpp.
  3. ref_shim.cpp compiles kernel.cl as C++ behind a macro shim next to the reference's host headers.
No reference source is copied into the repository; generated text only lands in the ignored _ref/.
"""
import os, re, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BZ2_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def sub_exact(text, old, new, count):
    found = text.count(old)
    if found != count:
        raise SystemExit(f"patch pattern {old!r}: expected {count} occurrence(s), found {found}")
    return text.replace(old, new)


def build(force=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping prebuilt oracle/_ref as is")
        return False
    os.makedirs(OUT, exist_ok=True)
    libs = [os.path.join(OUT, f"libref_{b}.so") for b in (10000, 100000)]
    shim = os.path.join(HERE, "ref_shim.cpp")
    if not force and all(os.path.exists(l) and os.path.getmtime(l) > os.path.getmtime(shim) and
                         os.path.getmtime(l) > os.path.getmtime(__file__) for l in libs):
        build_app()
        return True
    gen_cpp = os.path.join(OUT, "gen.cpp")
    with open(gen_cpp, "w") as f:
        f.write(f'#include "{REF}/kernel.cpp"\n'
                'int main(){ auto s=get_opencl_c_code(); fwrite(s.data(),1,s.size(),stdout); return 0; }\n')
    gen = os.path.join(OUT, "gen")
    run(["g++", "-std=c++17", "-w", f"-I{REF}", gen_cpp, "-o", gen])
    text = subprocess.check_output([gen]).decode()
    # P1
    text = sub_exact(text, "tableFrequencies[HUFFMAN_MAXIMUM_TABLES][HUFFMAN_MAXIMUM_ALPHABET_SIZE];",
                     "tableFrequencies[HUFFMAN_MAXIMUM_TABLES][HUFFMAN_MAXIMUM_ALPHABET_SIZE]={{0}};", 1)
    # P4
    text = sub_exact(text, "struct\nTRBudget\nbudget,\nint\nsize)", "struct\nTRBudget\n*budget,\nint\nsize)", 1)
    text = sub_exact(text, "updateTRBudget(&budget,", "updateTRBudget(budget,", 4)
    text = sub_exact(text, "last,\nbudget,\nn);", "last,\n&budget,\nn);", 1)
    cfg_src = open(os.path.join(REF, "include", "Config.hpp")).read()
    for base in (10000, 100000):
        d = os.path.join(OUT, f"cfg_{base}")
        os.makedirs(d, exist_ok=True)
        t = text
        c = cfg_src
        if base != 10000:  # P3
            t = sub_exact(t, "BLOCKSIZE_DEFAULT\n=\n10000;", f"BLOCKSIZE_DEFAULT\n=\n{base};", 1)
            c = sub_exact(c, "BLOCKSIZE_DEFAULT = 10000;", f"BLOCKSIZE_DEFAULT = {base};", 1)
        open(os.path.join(d, "kernel.cl"), "w").write(t)
        open(os.path.join(d, "Config.hpp"), "w").write(c)
        run(["g++", "-std=c++17", "-O2", "-w", "-fPIC", "-shared", "-pthread",
             f"-I{d}", f"-I{REF}/include", shim, "-o", os.path.join(OUT, f"libref_{base}.so")])
    build_app()
    return True


def build_app():
    """The reference's UNMODIFIED app.cpp compiled against the product's facade headers (bzip2_opencl_b200/include) and linked with
    libbz2b200.so: the drop-in claim, as a binary (oracle/_ref/app_b200_<base>) that the GPU tests run.  app.cpp includes
    "include/OutputStream.hpp" relative to its own directory, so it is fed to the compiler on stdin from a directory whose
    include/ is the facade; nothing of the reference is copied."""
    root = os.path.dirname(HERE)
    prod = os.path.join(root, "bzip2_opencl_b200")
    lib = os.path.join(prod, "libbz2b200.so")
    app = os.path.join(REF, "app.cpp")
    if not os.path.exists(lib) or not os.path.exists(app):
        return False
    shim = os.path.join(OUT, "facade_shim")
    os.makedirs(shim, exist_ok=True)
    link = os.path.join(shim, "include")
    if not os.path.islink(link):
        os.symlink(os.path.join(prod, "include"), link)
    inc = os.path.join(prod, "include")
    newest = max([os.path.getmtime(app), os.path.getmtime(os.path.join(root, "include", "bz2b200.h")), os.path.getmtime(__file__)] +
                 [os.path.getmtime(os.path.join(inc, f)) for f in os.listdir(inc)])
    for base in (10000, 100000):
        exe = os.path.join(OUT, f"app_b200_{base}")
        if os.path.exists(exe) and os.path.getmtime(exe) > newest:
            continue
        cmd = ["g++", "-std=c++17", "-O2", "-w", "-x", "c++", "-", "-o", exe, f"-DBZ2B200_BLOCKSIZE_DEFAULT={base}",
               "-I" + os.path.join(root, "include"), "-L" + prod, "-lbz2b200", "-Wl,-rpath," + prod, "-pthread"]
        print("+ (cd", shim, "&&", " ".join(cmd), "<", app + ")", flush=True)
        with open(app, "rb") as f:
            subprocess.check_call(cmd, stdin=f, cwd=shim)
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
