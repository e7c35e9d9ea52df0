"""Builds liblgmi.so (sm_100a) in-tree with nvcc + gcc.  No JIT cache: the .so
lives next to the sources so that it travels with a snapshot of the repo."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblgmi.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # the fp64 epilogue must not contract a*b+c
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".inl", ".c", ".h")):
                out.append(os.path.join(root, name))
    return out


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; liblgmi.so cannot be built")
    return nvcc


last_build_compiled = None     # True: the last build_lib() ran nvcc; False: it found the library newer than every source


def build_lib(force=False, verbose=False):
    """Compile csrc/ into l-giremi_b200/liblgmi.so.  Returns the path."""
    global last_build_compiled
    srcs = sources()
    if not force and _newer(LIB, srcs):
        last_build_compiled = False
        return LIB
    last_build_compiled = True
    nvcc = find_nvcc()
    obj = os.path.join(CSRC, "lgmi_lntab.o")
    cmd_c = ["gcc", "-O2", "-fPIC", "-c", os.path.join(CSRC, "lgmi_lntab.c"), "-o", obj]
    quad = subprocess.run(["gcc", "-print-file-name=libquadmath.a"], capture_output=True, text=True).stdout.strip()
    cmd_cu = [nvcc, *NVCC_FLAGS, "-shared", "-o", LIB,
              os.path.join(CSRC, "lgmi.cu"), obj, quad, "-Xlinker", "--exclude-libs,ALL"]
    for cmd in (cmd_c, cmd_cu):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("build failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
