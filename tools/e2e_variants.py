#!/usr/bin/env python
"""Dev timing of the pipelined host-to-host step on cfg2: input / output forms x number of groups.

    python tools/e2e_variants.py [chunks ...]        # default 4 6 8 12"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")
ctx = lg.Context(0)
pb = synth.make_uniform_planes(20261020, 20000, 50, 200, 0.5, chunk=500)


def pinned(a):
    p = ctx.pinned_empty(a.shape, a.dtype)
    p.array[...] = a
    return p


pin3, pin2, pint, pinf = pinned(pb.planes), pinned(pb.packed2()), pinned(pb.packed2(tight=True)), pinned(pb.site_flags)
forms = [("three planes, 16-byte rows", pin3, dict(), 0),
         ("two planes, split rows", pin2, dict(packed=True), lg.MODE_SPLIT_RECORDS),
         ("tight two planes, compact rows", pint, dict(tight=True), lg.MODE_COMPACT_OUTPUT)]
for chunks in [int(a) for a in sys.argv[1:]] or [4, 6, 8, 12]:
    pipe = lg.Pipeline(ctx, pb, chunks)
    for name, planes, kw, extra in forms:
        mode = lg.MODE_HET_ONLY | extra
        for _ in range(3):
            pipe.step(6, mode, planes.array, pinf.array, copy=False, **kw)
        t0 = time.perf_counter()
        for _ in range(10):
            r = pipe.step(6, mode, planes.array, pinf.array, copy=False, **kw)
        print("chunks=%2d %-32s %.3f ms  (%d rows)" % (chunks, name, (time.perf_counter() - t0) * 100, r.n_records), flush=True)
    pipe.close()
