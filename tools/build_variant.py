#!/usr/bin/env python
"""Kernel experiments: build another liblgmi.so with extra -D switches next to the product one.

    python tools/build_variant.py TAG -DLGMI_PHASE_CLOCKS ...   ->  build/liblgmi_TAG.so

Load it with LGMI_LIB=build/liblgmi_TAG.so (l-giremi_b200/_lib.py).  build/ is git-ignored and
travels to the GPU box with the snapshot."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bld = importlib.import_module("l-giremi_b200.build")


def main():
    tag, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liblgmi_%s.so" % tag)
    bld.build_lib()                                          # makes lgmi_lntab.o
    quad = subprocess.run(["gcc", "-print-file-name=libquadmath.a"], capture_output=True, text=True).stdout.strip()
    cmd = [bld.find_nvcc(), *bld.NVCC_FLAGS, *flags, "-shared", "-o", out, os.path.join(bld.CSRC, "lgmi.cu"),
           os.path.join(bld.CSRC, "lgmi_lntab.o"), quad, "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    for line in (r.stdout + r.stderr).split("\n"):
        if r.returncode or "k_pairs_fast" in line or "error" in line:
            print(line)
    if r.returncode:
        raise SystemExit("build failed")
    print(out)


if __name__ == "__main__":
    main()
