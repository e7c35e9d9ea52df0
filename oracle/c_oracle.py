"""ctypes wrapper + build recipe of oracle/oracle_mi.c.  TEST INFRASTRUCTURE ONLY
(see the header of oracle_mi.c); the product never imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_mi.c")
LIB = os.path.join(HERE, "liboracle_mi.so")

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        vp = C.c_void_p
        lib.oracle_mi_from_table.restype = C.c_double
        lib.oracle_mi_from_table.argtypes = [vp]
        lib.oracle_python_sum.restype = C.c_double
        lib.oracle_python_sum.argtypes = [vp, C.c_int64]
        lib.oracle_site_labels.restype = None
        lib.oracle_site_labels.argtypes = [C.c_int32, C.c_int32] + [vp] * 8
        lib.oracle_unit_pairs.restype = C.c_int64
        lib.oracle_unit_pairs.argtypes = [C.c_int32, C.c_int32, vp, vp, C.c_int32, vp, vp, vp, vp]
        lib.oracle_site_means.restype = None
        lib.oracle_site_means.argtypes = [C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp]
        lib.oracle_mip_calls.restype = None
        lib.oracle_mip_calls.argtypes = [C.c_int64, vp, vp, C.c_double, vp, vp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data


def mi_from_table(table) -> float:
    t = np.ascontiguousarray(np.asarray(table, dtype=np.int64).reshape(9))
    return float(load().oracle_mi_from_table(_p(t)))


def python_sum(values) -> float:
    v = np.ascontiguousarray(values, dtype=np.float64)
    return float(load().oracle_python_sum(_p(v), v.size))


TYPE_CODE = {"mismatch": 0, "snp": 1, "het_snp": 2}


def code_mismatches(mismatches):
    """Integer-code a reference ``mismatches`` dict, preserving every order the
    result depends on.  Returns a dict of arrays for `unit_step`."""
    positions = sorted(mismatches)
    read_id, allele_id = {}, {}
    ent_off, ent_allele, ent_read = [0], [], []
    dep_off, dep_allele, dep_count = [0], [], []
    for pos in positions:
        site = mismatches[pos]
        for allele, names in site['nt'].items():
            a = allele_id.setdefault(allele, len(allele_id))
            for name in names:
                ent_allele.append(a)
                ent_read.append(read_id.setdefault(name, len(read_id)))
        ent_off.append(len(ent_allele))
        for allele, depth in site['depth'].items():
            dep_allele.append(allele_id.setdefault(allele, len(allele_id)))
            dep_count.append(int(depth))
        dep_off.append(len(dep_allele))
    return dict(
        positions=positions,
        types=[mismatches[p]['type'] for p in positions],
        n_reads=len(read_id),
        ent_off=np.array(ent_off, np.int64), ent_allele=np.array(ent_allele, np.int32),
        ent_read=np.array(ent_read, np.int32), dep_off=np.array(dep_off, np.int64),
        dep_allele=np.array(dep_allele, np.int32), dep_count=np.array(dep_count, np.int64))


def labels_of(coded):
    S, R = len(coded['positions']), coded['n_reads']
    labels = np.empty((S, max(R, 0)), dtype=np.int8)
    bad = np.zeros(max(S, 1), dtype=np.uint8)
    load().oracle_site_labels(S, R, _p(coded['ent_off']), _p(coded['ent_allele']), _p(coded['ent_read']),
                              _p(coded['dep_off']), _p(coded['dep_allele']), _p(coded['dep_count']),
                              _p(labels), _p(bad))
    return labels, bad[:S]


def unit_pairs_from_labels(labels, bad, min_common):
    """labels int8 [S,R] (-1 uncovered).  Returns (i, j, mi, tables[n,9])."""
    labels = np.ascontiguousarray(labels, dtype=np.int8)
    S, R = labels.shape
    cap = max(1, S * (S - 1) // 2)
    oi, oj = np.empty(cap, np.int32), np.empty(cap, np.int32)
    omi, otab = np.empty(cap, np.float64), np.empty((cap, 9), np.int64)
    bad = np.ascontiguousarray(bad, dtype=np.uint8) if bad is not None else np.zeros(max(S, 1), np.uint8)
    n = load().oracle_unit_pairs(S, R, _p(labels), _p(bad), int(min_common), _p(oi), _p(oj), _p(omi), _p(otab))
    if n < 0:
        raise IndexError('list index out of range')
    return oi[:n].copy(), oj[:n].copy(), omi[:n].copy(), otab[:n].copy()


def site_means(n_sites, is_het, i, j, mi):
    is_het = np.ascontiguousarray(is_het, dtype=np.uint8)
    i, j = np.ascontiguousarray(i, np.int32), np.ascontiguousarray(j, np.int32)
    mi = np.ascontiguousarray(mi, np.float64)
    mean, cnt = np.empty(max(n_sites, 1), np.float64), np.empty(max(n_sites, 1), np.int32)
    load().oracle_site_means(n_sites, _p(is_het), len(mi), _p(i), _p(j), _p(mi), _p(mean), _p(cnt))
    return mean[:n_sites], cnt[:n_sites]


def unit_step(mismatches, min_common=5):
    """Whole per-unit step on a reference dict: returns dict(rows, mean)."""
    coded = code_mismatches(mismatches)
    labels, bad = labels_of(coded)
    i, j, mi, tab = unit_pairs_from_labels(labels, bad, min_common)
    pos, typ = coded['positions'], coded['types']
    rows = [[pos[a], typ[a], pos[b], typ[b], m] for a, b, m in zip(i.tolist(), j.tolist(), mi.tolist())]
    is_het = np.array([t == 'het_snp' for t in typ], dtype=np.uint8)
    mean, cnt = site_means(len(pos), is_het, i, j, mi)
    return dict(rows=rows, tables=tab, mean=mean, cnt=cnt, positions=pos, types=typ, i=i, j=j, mi=mi)


def mip_calls(mean, type_codes, threshold=0.05):
    mean = np.ascontiguousarray(mean, np.float64)
    t = np.ascontiguousarray(type_codes, np.uint8)
    mip, call = np.empty(mean.size, np.float64), np.empty(mean.size, np.uint8)
    load().oracle_mip_calls(mean.size, _p(mean), _p(t), float(threshold), _p(mip), _p(call))
    return mip, call
