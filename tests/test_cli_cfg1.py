"""BASELINE.json configs[0]: the reference's own `l-giremi` CLI (unmodified, from
baseline/_ref, pysam replaced by the stand-in over a simulated dataset) run twice --
stock on the CPU, and with this repository's MI step patched in (install(batched=True))
-- must write the same output tables."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import MI_ATOL, MI_RTOL, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import simdata  # noqa: E402

pytestmark = pytest.mark.gpu


def run_cli(tmp, prefix, patched, extra):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_cli.py")] + (["--patched"] if patched else []) + \
          [os.path.join(tmp, "ds.pkl"), os.path.join(tmp, prefix)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def same_table(a, b, float_cols=(), skip=()):
    ta, tb = pd.read_table(a), pd.read_table(b)
    assert list(ta.columns) == list(tb.columns) and len(ta) == len(tb), (a, ta.shape, tb.shape)
    for c in ta.columns:
        if c in skip:
            continue
        if c in float_cols:
            x, y = ta[c].to_numpy(dtype=float), tb[c].to_numpy(dtype=float)
            assert np.array_equal(np.isnan(x), np.isnan(y)), c
            ok = np.isnan(y) | (np.abs(x - y) <= MI_RTOL * np.abs(y) + MI_ATOL)
            assert ok.all(), (c, x[~ok][:3], y[~ok][:3])
        else:
            assert ta[c].equals(tb[c]), c
    return len(ta)


def test_cli_outputs_identical_stock_vs_patched(tmp_path, ref_giremi, gpu_ctx):
    tmp = str(tmp_path)
    ds = simdata.Dataset(seed=20261026, n_genes=12, reads_per_gene=200)
    with open(os.path.join(tmp, "ds.pkl"), "wb") as fh:
        pickle.dump(ds, fh)
    extra = ["-t", "2", "--mi_min_common_read", "6", "--mi_p_threshold", "0.05", "--min_total_depth", "2"]
    stock = run_cli(tmp, "stock", False, extra)
    patched = run_cli(tmp, "patched", True, extra)
    assert not stock["patched"] and patched["patched"]
    p = lambda prefix, ext: os.path.join(tmp, prefix + ext)
    n_mi = same_table(p("patched", ".mi.txt"), p("stock", ".mi.txt"), float_cols=("mi",))
    assert n_mi > 50
    same_table(p("patched", ".strand.txt"), p("stock", ".strand.txt"))
    same_table(p("patched", ".removed.txt"), p("stock", ".removed.txt"))
    # `score` comes from a model trained on an unseeded DataFrame.sample() (giremi.py:117-122): it differs
    # between two stock runs as well; everything the MI step feeds -- mean_mi, mip -- must agree
    n_sites = same_table(p("patched", ".mismatch.txt"), p("stock", ".mismatch.txt"),
                         float_cols=("mean_mi", "mip", "ratio", "allelic_ratio_diff"), skip=("score",))
    assert n_sites > 30
    # --mi_calculation_only: the three MI-step tables, nothing else
    only = run_cli(tmp, "only", True, extra + ["--mi_calculation_only"])
    assert only["patched"]
    same_table(p("only", ".mi.txt"), p("stock", ".mi.txt"), float_cols=("mi",))
    assert not os.path.exists(p("only", ".mismatch.txt"))
