#!/usr/bin/env python
"""Generates the golden fixtures in this directory by EXECUTING THE REAL
REFERENCE (imported from /root/reference, read-only) and the installed
scikit-learn.  Run in the build container only; the fixtures are committed,
the reference is not.

    python tests/golden/make_golden.py

Writes
    kat.json            known-answer vectors of SURVEY 8c, re-derived
    units_fuzz.json     random `mismatches` dicts + reference outputs
    units_synth.json    units from the package's synthetic generator + outputs
    tables.json         3x3 contingency tables + sklearn mutual_info_score
    ecdf.json           stat.ecdf / mip / threshold-call vectors
Floats are stored as C99 hex strings (bit-exact).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = "/root/reference/src/giremi"


def load_reference():
    """giremi/__init__.py needs an installed dist; bypass it with a stub package."""
    pkg = types.ModuleType("giremi")
    pkg.__path__ = [REF]
    pkg.__version__ = "0.2.4"
    sys.modules["giremi"] = pkg
    mi = importlib.import_module("giremi.mutual_information")
    stat = importlib.import_module("giremi.stat")
    return mi, stat


def hexf(x):
    return float(x).hex()


def run_unit(ref_mi, mismatches, min_common):
    """What mismatch.py:387-404 does for one strand, with the real functions."""
    full = ref_mi.mismatch_pair_mutual_info(mismatches, min_common_reads=min_common) if len(mismatches) > 1 else []
    kept = [a for a in full if a[1] == 'het_snp' or a[3] == 'het_snp']
    means = ref_mi.mean_mismatch_pair_mutual_info(kept) if kept else []
    return dict(
        min_common=min_common,
        mismatches={str(p): {'type': s['type'], 'ref': s['ref'], 'depth': s['depth'], 'nt': s['nt']}
                    for p, s in mismatches.items()},
        rows=[[r[0], r[1], r[2], r[3], hexf(r[4])] for r in full],
        kept_rows=[[r[0], r[1], r[2], r[3], hexf(r[4])] for r in kept],
        means=[[p, hexf(m)] for p, m in means],
    )


def site(pairs, typ='mismatch', ref='A', depth=None):
    nt = {a: list(names) for a, names in pairs}
    return {'ref': ref, 'type': typ, 'nt': nt, 'depth': depth or {a: len(v) for a, v in nt.items()},
            'neighbor': [], 'up': 'A', 'down': 'C'}


def r(lo, hi):
    return ['r%d' % k for k in range(lo, hi + 1)]


def make_kat(ref_mi, ref_stat):
    a = site([('G', r(0, 5)), ('A', r(6, 11))])
    b = site([('A', r(0, 5)), ('C', r(6, 11))], typ='het_snp')
    c = site([('G', r(0, 3)), ('A', r(4, 7)), ('T', r(8, 11))])
    d = site([('G', r(0, 5)), ('A', r(3, 11))])                      # r3-5 listed twice: last wins
    e = site([('G', r(0, 2)), ('A', r(6, 11))])
    f = site([('A', r(6, 11)), ('C', ['x1', 'x2', 'x3'])])
    cases = {
        'ab': ({10: a, 20: b}, 6), 'cb': ({10: c, 20: b}, 6), 'db': ({10: d, 20: b}, 6),
        'ef_min6': ({10: e, 20: f}, 6), 'ef_min7': ({10: e, 20: f}, 7),
        'abc': ({10: a, 20: b, 30: c}, 6),
    }
    out = {k: run_unit(ref_mi, m, mc) for k, (m, mc) in cases.items()}
    x = [0.1, 0.2, 0.2, 0.4]
    samples = [0.05, 0.1, 0.2, 0.3, 0.4, 0.5]
    fn = ref_stat.ecdf(x)
    out['ecdf'] = dict(x=x, samples=samples, y=[hexf(fn(s)) for s in samples])
    # one-allele depth -> IndexError once the pair survives
    bad = site([('G', r(0, 11))], depth={'G': 12})
    try:
        ref_mi.mismatch_pair_mutual_info({10: bad, 20: b}, 6)
        out['index_error'] = False
    except IndexError:
        out['index_error'] = True
    return out


def make_fuzz(ref_mi, n=90):
    from fuzz import random_mismatches
    rng = np.random.default_rng(20261018)
    units = []
    for k in range(n):
        if k % 9 == 8:      # a few larger ones: >= 8 non-zero cells, numpy's 8-accumulator sum
            m = random_mismatches(rng, n_sites=int(rng.integers(5, 9)), n_reads=int(rng.integers(60, 160)),
                                  cov=0.9, p_multi=0.9)
        else:
            m = random_mismatches(rng)
        units.append(run_unit(ref_mi, m, int(rng.choice([1, 2, 3, 5, 6, 8]))))
    return units


def make_synth(ref_mi):
    lg = importlib.import_module("l-giremi_b200.synth")
    out = []
    for seed, G, S, R, cov, mc in [(11, 3, 12, 48, 0.5, 6), (12, 2, 20, 130, 0.7, 6), (13, 2, 9, 300, 0.4, 10)]:
        sb = lg.make_uniform(seed, G, S, R, cov)
        for g in range(G):
            u = run_unit(ref_mi, sb.mismatches(g), mc)
            u['synth'] = dict(seed=seed, G=G, S=S, R=R, cov=cov, g=g)
            out.append(u)
    return out


def make_tables():
    from sklearn.metrics import mutual_info_score
    rng = np.random.default_rng(7)
    out = []

    def add(t):
        t = np.asarray(t, dtype=np.int64).reshape(3, 3)
        if t.sum() == 0:
            return
        l1 = [a for a in range(3) for b in range(3) for _ in range(int(t[a, b]))]
        l2 = [b for a in range(3) for b in range(3) for _ in range(int(t[a, b]))]
        out.append(dict(table=t.reshape(-1).tolist(), mi=hexf(mutual_info_score(l1, l2))))

    for _ in range(1500):
        scale = int(rng.choice([3, 10, 40, 200, 3000]))
        t = rng.integers(0, scale, size=9)
        mask = rng.random(9) < rng.choice([0.0, 0.3, 0.6])
        t[mask] = 0
        add(t)
    for _ in range(600):                                            # exact independence (MI == 0 analytically)
        ra, cb = rng.integers(0, 12, 3), rng.integers(0, 12, 3)
        add(np.outer(ra, cb))
    for _ in range(300):                                            # 2x2 in the minor/major block
        t = np.zeros(9, dtype=np.int64)
        t[[4, 5, 7, 8]] = rng.integers(0, int(rng.choice([4, 30, 250])), 4)
        add(t)
    for n in (1, 2, 5, 6, 7, 100, 100000):                          # degenerate / single class
        add([0, 0, 0, 0, 0, 0, 0, 0, n])
        add([0, 0, 0, 0, n, n, 0, 0, 0])
        add([0, 0, 0, 0, n, 0, 0, 0, n])
    add([30000, 1, 2, 3, 40000, 5, 6, 7, 29999])                    # deep unit scale
    return out


def make_ecdf(ref_stat):
    rng = np.random.default_rng(99)
    out = []
    for n in (1, 2, 3, 7, 50, 1000):
        x = np.round(rng.random(n), int(rng.choice([1, 2, 12])))    # ties at coarse rounding
        fn = ref_stat.ecdf(x)
        samples = np.concatenate([x[: min(n, 20)], rng.random(20), [0.0, 1.0, -1.0, 2.0]])
        out.append(dict(x=[hexf(v) for v in x], samples=[hexf(v) for v in samples],
                        y=[hexf(fn(s)) for s in samples]))
    # the CLI's global pass (giremi.py:415-429, 97-114) on a synthetic site table
    n = 400
    mean = rng.random(n).round(2)
    mean[rng.random(n) < 0.25] = np.nan
    typ = rng.choice(['mismatch', 'snp', 'het_snp'], size=n, p=[0.7, 0.1, 0.2])
    het = [m for m, t in zip(mean, typ) if t == 'het_snp' and not np.isnan(m)]
    fn = ref_stat.ecdf(het)
    mip = [fn(m) if not np.isnan(m) else np.nan for m in mean]
    thr = 0.05
    pos = [(not np.isnan(m)) and p <= thr and t == 'mismatch' for m, p, t in zip(mean, mip, typ)]
    neg = [(not np.isnan(m)) and p > thr and t != 'mismatch' for m, p, t in zip(mean, mip, typ)]
    table = dict(mean=[hexf(v) for v in mean], type=typ.tolist(), mip=[hexf(v) for v in mip],
                 threshold=thr, positive=[bool(v) for v in pos], negative=[bool(v) for v in neg])
    return dict(functions=out, site_table=table)


def main():
    t0 = time.time()
    ref_mi, ref_stat = load_reference()
    import sklearn
    meta = dict(generated_by="tests/golden/make_golden.py", reference="gxiaolab/L-GIREMI v0.2.4 (/root/reference)",
                sklearn=sklearn.__version__, numpy=np.__version__, python=sys.version.split()[0])
    outputs = {
        "kat.json": make_kat(ref_mi, ref_stat),
        "units_fuzz.json": make_fuzz(ref_mi),
        "units_synth.json": make_synth(ref_mi),
        "tables.json": make_tables(),
        "ecdf.json": make_ecdf(ref_stat),
    }
    for name, payload in outputs.items():
        with open(os.path.join(HERE, name), "w") as fh:
            json.dump(dict(meta=meta, data=payload), fh, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")
    # the timing port must be row-identical to the reference (it stands in for it on the GPU box)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port
    from fuzz import random_mismatches
    rng = np.random.default_rng(5)
    for _ in range(40):
        m = random_mismatches(rng)
        a = ref_mi.mismatch_pair_mutual_info(m, 3)
        b = ref_port.port_pair_mutual_info(m, 3)
        assert a == b, "port diverges from the reference"
        ka = [x for x in a if x[1] == 'het_snp' or x[3] == 'het_snp']
        assert ref_mi.mean_mismatch_pair_mutual_info(ka) == ref_port.port_mean_mutual_info(ka)
    lg = importlib.import_module("l-giremi_b200.synth")
    m = lg.make_uniform(3, 1, 50, 200, 0.5).mismatches(0)
    t1 = time.time(); a = ref_mi.mismatch_pair_mutual_info(m, 6); t2 = time.time()
    b = ref_port.port_pair_mutual_info(m, 6); t3 = time.time()
    assert a == b
    print("reference %.3fs  port %.3fs on one S=50 R=200 unit (%d rows)" % (t2 - t1, t3 - t2, len(a)))
    print("done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
