"""Timing port of the reference's CPU MI step.  TEST/BENCH INFRASTRUCTURE ONLY.

The real reference (a Python package) cannot travel to the GPU box, so the
``cpu_baseline`` / ``--impl reference`` legs of bench.py time this port
instead.  It performs the same operations with the same libraries as
/root/reference/src/giremi/mutual_information.py:6-60 -- per pair it rebuilds
both read->allele dicts, intersects and sorts the read names, ranks alleles by
depth and calls scikit-learn's ``mutual_info_score`` on two Python lists -- so
its cost profile is the reference's (validated against the real functions in
tests/golden/make_golden.py: identical rows, same wall time within noise).
It is written from the algorithm's description, not copied."""
from __future__ import annotations

from itertools import combinations

from sklearn.metrics import mutual_info_score


def _read_to_allele(site):
    table = {}
    for allele, names in site['nt'].items():
        for name in names:
            table[name] = allele
    return table


def _allele_codes(site):
    ranked = sorted(site['depth'].items(), key=lambda kv: kv[1], reverse=True)
    return {ranked[1][0]: 1, ranked[0][0]: 2}


def port_pair_mutual_info(mismatches, min_common_reads=5):
    rows = []
    for pa, pb in combinations(sorted(mismatches), 2):
        sa, sb = mismatches[pa], mismatches[pb]
        ra, rb = _read_to_allele(sa), _read_to_allele(sb)
        shared = sorted(name for name in ra if name in rb)
        if len(shared) < min_common_reads:
            continue
        ca, cb = _allele_codes(sa), _allele_codes(sb)
        la = [ca.get(ra[name], 0) for name in shared]
        lb = [cb.get(rb[name], 0) for name in shared]
        rows.append([pa, sa['type'], pb, sb['type'], mutual_info_score(la, lb)])
    return rows


def port_mean_mutual_info(rows):
    groups = {}
    for pa, _ta, pb, _tb, mi in rows:
        groups.setdefault(pa, []).append(mi)
        groups.setdefault(pb, []).append(mi)
    return [[pos, sum(v) / len(v)] for pos, v in groups.items()]


def port_unit_step(mismatches, min_common_reads=5):
    """mismatch.py:387-404 for one strand: pairs -> het filter -> means."""
    rows = port_pair_mutual_info(mismatches, min_common_reads) if len(mismatches) > 1 else []
    kept = [r for r in rows if r[1] == 'het_snp' or r[3] == 'het_snp']
    return rows, kept, (port_mean_mutual_info(kept) if kept else [])


def port_chunk(args):
    """Pool worker: a chunk of units (giremi.py:367-380 chunks footprints the same way)."""
    units, min_common = args
    n_pairs = 0
    out = []
    for m in units:
        rows, kept, means = port_unit_step(m, min_common)
        n = len(m)
        n_pairs += n * (n - 1) // 2
        out.append((len(rows), len(kept), len(means)))
    return n_pairs, out


# --------------------------------------------------------------------------- the real thing, when it is installed
def reference_functions():
    """(mismatch_pair_mutual_info, mean_mismatch_pair_mutual_info) of the UNMODIFIED reference,
    pip-installed under baseline/_ref by __graft_entry__.build(); None when it is absent."""
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.isdir(os.path.join(path, "giremi")):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        from giremi.mutual_information import mean_mismatch_pair_mutual_info, mismatch_pair_mutual_info
    except Exception:                                         # noqa: BLE001
        return None
    return mismatch_pair_mutual_info, mean_mismatch_pair_mutual_info


def reference_chunk(args):
    """Pool worker running the reference's own two functions the way mismatch.py:387-404 does."""
    units, min_common = args
    pair_mi, mean_mi = reference_functions()
    n_pairs = 0
    out = []
    for m in units:
        rows = pair_mi(m, min_common_reads=min_common) if len(m) > 1 else []
        kept = [r for r in rows if r[1] == 'het_snp' or r[3] == 'het_snp']
        means = mean_mi(kept) if kept else []
        n = len(m)
        n_pairs += n * (n - 1) // 2
        out.append((len(rows), len(kept), len(means)))
    return n_pairs, out
