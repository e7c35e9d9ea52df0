#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: mi_min_common_read x site-coverage sweep over
20 000 units x 50 sites x 200 reads.  Device-resident whole-step time per grid point
(best of 5 after 2 warm-ups) and the surviving fraction.  Dev measurement, not a bench line."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")

G = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = lg.Context(0)
rows = []
for cov in (0.01, 0.05, 0.1, 0.25, 0.5):
    pb = synth.make_uniform_planes(20261030 + int(cov * 100), G, 50, 200, cov, chunk=500)
    b = lg.Batch(ctx, pb)
    b.upload()
    for mc in (6, 10, 20, 50):
        ms = []
        for k in range(7):
            b.run(mc, lg.MODE_ALL_PAIRS)
            r = b.sync()
            if k >= 2:
                ms.append(float(r.kernel_ms))
        rows.append({"cov": cov, "min_common": mc, "step_ms": min(ms), "pairs_per_s": pb.n_candidates / (min(ms) * 1e-3),
                     "surviving_fraction": int(r.n_records) / pb.n_candidates})
        print(json.dumps(rows[-1]), flush=True)
    b.close()
json.dump({"workload": "cfg5 grid on one B200: %d units x 50 sites x 200 reads, ALL_PAIRS" % G, "grid": rows},
          open(os.path.join(ROOT, "gpurun_out", "cfg5_sweep.json"), "w"), indent=1)
