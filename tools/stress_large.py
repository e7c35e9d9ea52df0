#!/usr/bin/env python
"""One-off robustness run: 120 000 units (cfg2 replicated 6x, 147 M candidate pairs, 0.58 GB of
planes) through one submit and through the pipelined step; every replica must reproduce the
first copy's rows."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")
enc = importlib.import_module("l-giremi_b200.encode")

base = synth.make_uniform_planes(20261020, 20000, 50, 200, 0.5, chunk=500)
K = 6
big = enc.concat_plane_batches([base] * K)
ctx = lg.Context(0)
t0 = time.time()
one = lg.mi_step_batched(big, 6, lg.MODE_HET_ONLY, ctx=ctx, n_chunks=1)
t1 = time.time()
pipe = lg.Pipeline(ctx, big, 8)
res = pipe.step(6, lg.MODE_HET_ONLY | lg.MODE_SPLIT_RECORDS, packed=True)
t2 = time.time()
assert one.n_candidates == 24_500_000 * K and res.n_records == one.n_records
assert np.array_equal(res.records, one.records) and np.array_equal(res.site_mean, one.site_mean, equal_nan=True)
n = one.n_records // K
first = one.records[:n]
for k in range(1, K):
    part = one.records[k * n:(k + 1) * n]
    assert np.array_equal(part['mi'], first['mi']) and np.array_equal(part['i'], first['i'])
    assert np.array_equal(part['unit'], first['unit'] + 20000 * k)
print("stress ok: %d units, %d candidate pairs, %d rows; one submit %.2f s, pipelined %.2f s (incl. setup)"
      % (big.n_units, one.n_candidates, one.n_records, t1 - t0, t2 - t1))
