#!/usr/bin/env python
"""Where a k_pairs_fast CTA spends its cycles: per barrier of the unit loop, the cycles each warp
worked before arriving and the cycles it then waited (debug build, -DLGMI_PHASE_CLOCKS).

    python tools/build_variant.py clk -DLGMI_PHASE_CLOCKS
    LGMI_LIB=build/liblgmi_clk.so python tools/phase_clocks.py        # on the GPU box

Workload: bench.py's cfg2 batch.  Not a timing run (clock64 + atomics perturb the kernel)."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lg = importlib.import_module("l-giremi_b200")
synth = importlib.import_module("l-giremi_b200.synth")
_lib = importlib.import_module("l-giremi_b200._lib")

PHASES = ["prefetch+zero -> B1 (rows landed)", "land -> B2", "site lists + other cells -> B3", "counts -> B4",
          "prefix + MI -> B5", "means / emit -> B6 (end of unit)"]

pb = synth.make_uniform_planes(20261020, 20000, 50, 200, 0.5, chunk=500)
ctx = lg.Context(0)
b = lg.Batch(ctx, pb)
b.upload()
lib = C.CDLL(_lib.LIB_PATH)
buf = np.zeros((8, 8, 2), dtype=np.uint64)
b.run(6, lg.MODE_ALL_PAIRS)
b.sync()
lib.lgmi_debug_phase_clocks(None, 1)
b.run(6, lg.MODE_ALL_PAIRS)
r = b.sync()
lib.lgmi_debug_phase_clocks(buf.ctypes.data_as(C.c_void_p), 0)
per_unit = buf.astype(np.float64) / pb.n_units
out = {"pairs_kernel_ms_instrumented": float(r.pairs_kernel_ms), "units": pb.n_units, "cycles_per_unit": {}}
for k, name in enumerate(PHASES):
    out["cycles_per_unit"][name] = {"work_by_warp": [round(x) for x in per_unit[k, :, 0]],
                                    "wait_by_warp": [round(x) for x in per_unit[k, :, 1]]}
out["total_cycles_per_unit_warp0"] = round(float(per_unit[:, 0, :].sum()))
print(json.dumps(out, indent=1))
