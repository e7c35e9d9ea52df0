#!/usr/bin/env python
"""Where the host spends the streamed end-to-end step: seconds inside lgmi_pipeline_begin_packed /
lgmi_pipeline_collect / lgmi_pipeline_finish per cfg2 step, three steps in flight, for a few group counts.

    python tools/stream_host_times.py [steps [GROUPSxDEPTH ...]]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    ctx = lg.get_context(0)
    pb = synth.make_uniform_planes(20261020, 20000, 50, 200, 0.5)
    packed = pb.packed2(tight=True)
    pin_p = ctx.pinned_empty(packed.shape, np.uint32)
    pin_p.array[...] = packed
    pin_f = ctx.pinned_empty(pb.site_flags.shape, np.uint8)
    pin_f.array[...] = pb.site_flags
    mode = lg.MODE_HET_ONLY | lg.MODE_COMPACT_OUTPUT
    combos = [(int(a), int(b)) for a, b in (x.split("x") for x in sys.argv[2:])] or [(2, 3), (3, 3), (4, 3), (6, 3)]
    for groups, depth in combos:                              # "groups x steps in flight"
        pipes = [lg.Pipeline(ctx, pb, groups) for _ in range(depth)]
        for p in pipes:
            p.step(6, mode, pin_p.array, pin_f.array, copy=False, tight=True)
        tb = tc = tf = 0.0
        t0 = time.perf_counter()
        for k in range(steps):
            a = time.perf_counter()
            pipes[k % depth].begin(6, mode, pin_p.array, pin_f.array, tight=True)
            b = time.perf_counter()
            if depth > 2 and k >= depth - 2:
                pipes[(k - depth + 2) % depth].collect()
            c = time.perf_counter()
            if k >= depth - 1:
                pipes[(k - depth + 1) % depth].finish(copy=False)
            d = time.perf_counter()
            tb, tc, tf = tb + (b - a), tc + (c - b), tf + (d - c)
        for k in range(max(0, steps - depth + 1), steps):
            pipes[k % depth].finish(copy=False)
        total = time.perf_counter() - t0
        print(json.dumps({"groups": groups, "in_flight": depth, "ms_per_step": round(1e3 * total / steps, 3),
                          "host_ms_in_begin": round(1e3 * tb / steps, 3), "host_ms_in_collect": round(1e3 * tc / steps, 3),
                          "host_ms_in_finish": round(1e3 * tf / steps, 3)}), flush=True)
        for p in pipes:
            p.close()


if __name__ == "__main__":
    main()
