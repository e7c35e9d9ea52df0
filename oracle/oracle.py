"""CPU oracle for the L-GIREMI mutual-information step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``l-giremi_b200``) never does: it fails loudly without its CUDA library.

This is a from-scratch restatement (not a copy) of the reference algorithm:

  * pair enumeration, common-read filter, allele->label map
        /root/reference/src/giremi/mutual_information.py:6-45
  * per-site mean MI
        /root/reference/src/giremi/mutual_information.py:48-60
  * het filter between the two
        /root/reference/src/giremi/mismatch.py:393-396
  * empirical CDF / ``mip``
        /root/reference/src/giremi/stat.py:7-29,
        /root/reference/src/giremi/script/giremi.py:415-429
  * threshold call
        /root/reference/src/giremi/script/giremi.py:97-114

The MI arithmetic itself lives in a third-party dependency the reference does
not vendor: scikit-learn ``mutual_info_score`` (pyproject.toml lists it
unpinned; 1.9.0 is installed in this image) --
``sklearn/metrics/cluster/_supervised.py:822-935`` with the contingency matrix
from ``:96-182``.  ``mi_from_table`` restates that published algorithm
including its term association, eps-zeroing, 1-class shortcut, clip, and
numpy's pairwise summation order.

Parity is PINNED: ``tests/golden/make_golden.py`` executed the real reference
(imported from /root/reference) and the installed scikit-learn and wrote the
fixtures under ``tests/golden/``; ``tests/test_oracle.py`` checks this module
against them bit-for-bit.

Two entry levels are provided:

  * dict level  -- operates on the reference's own ``mismatches[strand]`` dict
    (used to pin the oracle against the reference),
  * table level -- ``mi_from_table`` on a 3x3 integer table with label order
    (other, minor, major), which is what the CUDA kernels are checked against.
"""
from __future__ import annotations

import math
from itertools import combinations

import numpy as np

_EPS = float(np.finfo(np.float64).eps)

OTHER, MINOR, MAJOR = 0, 1, 2  # label values == table index (mutual_information.py:33-38)


# --------------------------------------------------------------------------- #
# numpy / CPython arithmetic restated
# --------------------------------------------------------------------------- #
def numpy_sum_order(values):
    """Sum <=128 doubles the way ``ndarray.sum()`` does for a contiguous 1-D
    float64 array (numpy pairwise_sum: plain loop below 8 elements, otherwise
    8 running accumulators combined as a balanced tree, then the tail)."""
    n = len(values)
    if n < 8:
        acc = 0.0
        for v in values:
            acc += v
        return acc
    assert n <= 128
    r = [float(v) for v in values[:8]]
    i = 8
    stop = n - (n % 8)
    while i < stop:
        for k in range(8):
            r[k] += values[i + k]
        i += 8
    acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        acc += values[i]
        i += 1
    return acc


def python_sum_order(values):
    """``builtins.sum`` over floats as CPython >= 3.12 computes it
    (Neumaier-compensated; the first element is added to the int 0)."""
    it = iter(values)
    try:
        s = 0 + next(it)
    except StopIteration:
        return 0
    c = 0.0
    for x in it:
        t = s + x
        if abs(s) >= abs(x):
            c += (s - t) + x
        else:
            c += (x - t) + s
        s = t
    if c and math.isfinite(c):
        s += c
    return s


# --------------------------------------------------------------------------- #
# scikit-learn mutual_info_score restated on a contingency table
# --------------------------------------------------------------------------- #
def mi_from_table(table) -> float:
    """MI (nats) of a contingency table, bit-compatible with
    ``sklearn.metrics.mutual_info_score`` (_supervised.py:822-935) on the
    labels that produced it.  Rows/columns that are entirely zero correspond
    to labels that do not occur and are dropped (``np.unique`` in
    ``contingency_matrix``, :162-163)."""
    t = [[int(v) for v in row] for row in table]
    nr, nc = len(t), len(t[0])
    row_sum = [sum(t[i]) for i in range(nr)]
    col_sum = [sum(t[i][j] for i in range(nr)) for j in range(nc)]
    rows = [i for i in range(nr) if row_sum[i] > 0]
    cols = [j for j in range(nc) if col_sum[j] > 0]
    if len(rows) <= 1 or len(cols) <= 1:      # :920  "pi.size == 1 or pj.size == 1"
        return 0.0
    n_total = sum(row_sum)
    ln_total = math.log(n_total)              # math.log on the grand total (:927-929)
    terms = []
    for i in rows:                            # sp.find order: row-major over present classes
        for j in cols:
            n = t[i][j]
            if n == 0:
                continue
            q = n / n_total                                   # :924
            ln_n = float(np.log(np.int64(n)))                 # :923 (numpy log)
            outer = row_sum[i] * col_sum[j]                   # :926 int64 product
            ln_outer = -float(np.log(np.int64(outer))) + ln_total + ln_total   # :929
            term = q * (ln_n - ln_total) + q * ln_outer       # :930-933
            if abs(term) < _EPS:                              # :934
                term = 0.0
            terms.append(term)
    total = numpy_sum_order(terms)
    return float(total) if total > 0.0 else 0.0               # :935 clip(lower=0)


# --------------------------------------------------------------------------- #
# dict-level restatement of mutual_information.py
# --------------------------------------------------------------------------- #
def site_read_alleles(site) -> dict:
    """read name -> allele for one site; a name listed twice keeps the LAST
    allele in ('nt' key order, list order)  (mutual_information.py:15-16)."""
    out = {}
    for allele, names in site['nt'].items():
        for name in names:
            out[name] = allele
    return out


def site_major_minor(site):
    """(major, minor) allele by site-wide ``depth``, descending, ties keeping
    dict order (stable sort)  (mutual_information.py:25-32).  Raises
    IndexError when fewer than two alleles have a depth, as the reference
    does at :30."""
    ranked = sorted(site['depth'].items(), key=lambda kv: -kv[1])
    return ranked[0][0], ranked[1][0]


def site_labels(site) -> dict:
    """read name -> label in {0 other, 1 minor, 2 major}."""
    major, minor = site_major_minor(site)
    lab = {}
    for name, allele in site_read_alleles(site).items():
        lab[name] = MAJOR if allele == major else (MINOR if allele == minor else OTHER)
    return lab


def pair_table(site1, site2):
    """3x3 contingency table over the reads two sites share, index = label."""
    l1, l2 = site_labels(site1), site_labels(site2)
    table = [[0, 0, 0], [0, 0, 0], [0, 0, 0]]
    for name, a in l1.items():
        b = l2.get(name)
        if b is not None:
            table[a][b] += 1
    return table


def pair_mutual_info(mismatches, min_common_reads=5, with_tables=False):
    """Restates ``mismatch_pair_mutual_info``: rows ``[p1, type1, p2, type2, mi]``
    for every position pair (ascending, lexicographic) sharing at least
    ``min_common_reads`` de-duplicated reads (strict ``<`` drop, :19)."""
    rows = []
    tables = []
    positions = sorted(mismatches)
    reads = {p: site_read_alleles(mismatches[p]) for p in positions}
    for p1, p2 in combinations(positions, 2):
        r1, r2 = reads[p1], reads[p2]
        n_common = sum(1 for name in r1 if name in r2)
        if n_common < min_common_reads:
            continue
        table = pair_table(mismatches[p1], mismatches[p2])
        rows.append([p1, mismatches[p1]['type'], p2, mismatches[p2]['type'],
                     mi_from_table(table)])
        tables.append(table)
    return (rows, tables) if with_tables else rows


def het_filter(rows):
    """mismatch.py:393-396 -- keep a pair iff either site is a het SNP."""
    return [r for r in rows if r[1] == 'het_snp' or r[3] == 'het_snp']


def mean_pair_mutual_info(rows):
    """Restates ``mean_mismatch_pair_mutual_info``: ``[[pos, mean], ...]`` in
    first-appearance order; each mean is ``sum(values)/len(values)`` with the
    values in row order and CPython's float ``sum``."""
    per_site = {}
    for p1, _t1, p2, _t2, mi in rows:
        per_site.setdefault(p1, []).append(mi)
        per_site.setdefault(p2, []).append(mi)
    return [[pos, python_sum_order(v) / len(v)] for pos, v in per_site.items()]


def mi_step(mismatches, min_common_reads=5):
    """The whole per-unit step as ``region_mismatch_analysis`` runs it
    (mismatch.py:387-404): all pairs -> het filter -> per-site mean."""
    full = pair_mutual_info(mismatches, min_common_reads) if len(mismatches) > 1 else []
    kept = het_filter(full)
    means = mean_pair_mutual_info(kept) if kept else []
    return full, kept, means


# --------------------------------------------------------------------------- #
# global pass: ecdf, mip, threshold call
# --------------------------------------------------------------------------- #
def linspace_restated(start, stop, num):
    """numpy.linspace(start, stop, num) for num >= 1, endpoint=True."""
    div = num - 1
    y = np.arange(0, num, dtype=np.float64)
    delta = stop - start
    if div > 0:
        step = delta / div
        if step == 0:
            y = y / div
            y = y * delta
        else:
            y = y * step
    else:
        y = y * delta
    y = y + start
    if num > 1:
        y[-1] = stop
    return y


def ecdf_table(het_means):
    """(sorted x, y) of stat.py:16-19."""
    x = np.sort(np.asarray(het_means, dtype=np.float64))
    n = len(x)
    y = np.concatenate([[0.0], linspace_restated(1 / n, 1.0, n)])
    return x, y


def mip_values(mean_mi, is_het):
    """giremi.py:415-429: NaN stays NaN, otherwise the fraction of het-SNP
    means strictly below the site's mean (searchsorted side='left')."""
    mean_mi = np.asarray(mean_mi, dtype=np.float64)
    is_het = np.asarray(is_het, dtype=bool)
    out = np.full(mean_mi.shape, np.nan)
    valid = ~np.isnan(mean_mi)
    if valid.sum() == 0:
        return out
    x, y = ecdf_table(mean_mi[valid & is_het])
    out[valid] = y[np.searchsorted(x, mean_mi[valid], side='left')]
    return out


CALL_NONE, CALL_POS, CALL_NEG = 0, 1, 2


def threshold_calls(mean_mi, mip, is_mismatch, threshold=0.05):
    """giremi.py:97-114: 1 = positive training label (mismatch, mip<=thr),
    2 = negative (non-mismatch, mip>thr), 0 = neither."""
    mean_mi = np.asarray(mean_mi, dtype=np.float64)
    mip = np.asarray(mip, dtype=np.float64)
    is_mismatch = np.asarray(is_mismatch, dtype=bool)
    valid = ~np.isnan(mean_mi)
    call = np.zeros(mean_mi.shape, dtype=np.uint8)
    with np.errstate(invalid='ignore'):
        call[valid & (mip <= threshold) & is_mismatch] = CALL_POS
        call[valid & (mip > threshold) & ~is_mismatch] = CALL_NEG
    return call


# --------------------------------------------------------------------------- #
# site x splice-site MI (script/calculate_site_splice_mi.py:106-125)
# --------------------------------------------------------------------------- #
def site_splice_mutual_info(sites, splices, pairs):
    """For every (site label, allele, splice label) of `pairs`: over ALL list entries of the
    site's reads, sorted (:117-120; a read listed twice counts twice), the 0/1 vectors
    "read carries the allele" and "read has the splice site" (:121-122, membership by NAME),
    and their mutual_info_score (:123).  `sites`: {label: {allele: [read names]}},
    `splices`: {label: [read names]}."""
    out = []
    for site_label, seq, splice_label in pairs:
        with_seq_names = set(sites[site_label][seq])
        with_splice_names = set(splices.get(splice_label, ()))
        reads = sorted(name for allele in sites[site_label] for name in sites[site_label][allele])
        table = [[0, 0], [0, 0]]
        for name in reads:
            table[int(name in with_seq_names)][int(name in with_splice_names)] += 1
        out.append(mi_from_table(table))
    return out
