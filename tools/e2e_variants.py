#!/usr/bin/env python
"""Dev timing of the pipelined host-to-host step on cfg2 in its four input/output forms."""
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
pin3 = ctx.pinned_empty(pb.planes.shape, np.uint32); pin3.array[...] = pb.planes
p2 = pb.packed2()
pin2 = ctx.pinned_empty(p2.shape, np.uint32); pin2.array[...] = p2
pinf = ctx.pinned_empty(pb.site_flags.shape, np.uint8); pinf.array[...] = pb.site_flags
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pipe = lg.Pipeline(ctx, pb, chunks)
for packed in (False, True):
    for split in (False, True):
        mode = lg.MODE_HET_ONLY | (lg.MODE_SPLIT_RECORDS if split else 0)
        planes = pin2.array if packed else pin3.array
        for _ in range(3):
            pipe.step(6, mode, planes, pinf.array, copy=False, packed=packed)
        t0 = time.perf_counter()
        for _ in range(10):
            pipe.step(6, mode, planes, pinf.array, copy=False, packed=packed)
        print("packed=%d split=%d chunks=%d: %.3f ms" % (packed, split, chunks, (time.perf_counter() - t0) * 100))
