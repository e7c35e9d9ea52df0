#!/usr/bin/env python
"""Dev timing of the heavy-tailed workload (BASELINE.json configs[3] shape, fewer units):
whole-step device time and the share of each kernel group.  Not a bench line."""
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

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
t0 = time.time()
pb, _ = synth.make_heavy_tail(20261023, G)
gen_s = time.time() - t0
ctx = lg.Context(0)
b = lg.Batch(ctx, pb)
b.upload()
S = pb.units['n_sites'].astype(np.int64)
R = pb.units['n_reads'].astype(np.int64)
small = (S <= 64) & (R <= 256)
pairs = S * (S - 1) // 2
out = []
for _ in range(6):
    b.run(6, lg.MODE_ALL_PAIRS)
    r = b.sync()
    out.append((float(r.kernel_ms), float(r.pairs_kernel_ms), float(r.dense_kernel_ms)))
k_ms, fast_ms, dense_ms = min(out)
print(json.dumps({
    "units": G, "gen_s": round(gen_s, 1), "pairs": int(pairs.sum()), "pairs_small_units": int(pairs[small].sum()),
    "units_small": int(small.sum()), "word_pairs": int((pairs * ((R + 31) // 32)).sum()),
    "n_dense_units": int(r.n_dense_units), "step_ms": k_ms, "k_pairs_fast_ms": fast_ms, "dense_ms": dense_ms,
    "pairs_per_s": float(pairs.sum()) / (k_ms * 1e-3), "records": int(r.n_records)}))
