"""The fp64 arithmetic the kernels run (l-giremi_b200/csrc/lgmi_math.cuh),
compiled for the host, against the sklearn golden tables and the oracle.
The device compiles the same header with explicitly rounded intrinsics
(__dadd_rn ...), so bit-equality here carries over; tests/test_gpu_parity.py
re-checks it on the GPU."""
import math
import os
import sys

import numpy as np

from conftest import ROOT, assert_mi_close, unhex

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import c_oracle  # noqa: E402
import oracle  # noqa: E402


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def test_lntab_hi_is_rounded_log(lntab):
    """hi = RN(ln k) from binary128.  glibc's and numpy's log are faithfully but
    not correctly rounded: on this image they differ from RN at ~2e-5 resp.
    ~4e-5 of integer arguments, always by one ulp (and numpy's depends on the
    CPU's SIMD dispatch).  The table takes the platform-independent value; a
    one-ulp difference in a log moves MI by ~1e-16 relative, far inside the
    1e-10 tolerance of the MI parity tests."""
    k = np.arange(1, 1 << 17)
    hi = lntab[0::2][1:]
    lo = lntab[1::2][1:]
    for ref in (np.log(k.astype(np.float64)), np.array([math.log(int(x)) for x in k])):
        diff = hi != ref
        assert diff.mean() < 1e-4
        assert np.all(np.abs(hi[diff] - ref[diff]) <= np.spacing(hi[diff]))
    assert np.all(np.abs(lo[1:]) <= np.spacing(np.abs(hi[1:])) / 2)
    assert lntab[2] == 0.0 and lntab[3] == 0.0        # ln 1


def test_ln_product_is_rounded_log_of_product(math_host, lntab):
    rng = np.random.default_rng(3)
    a = rng.integers(1, 1 << 17, 20000)
    b = rng.integers(1, 1 << 17, 20000)
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    n_off = 0
    for x, y in zip(a.tolist(), b.tolist()):
        got = math_host.t_ln_product(x, y, lntab.ctypes.data)
        ref = float(np.log(np.int64(x * y)))
        if got != ref:                       # numpy not correctly rounded here: ours must be
            n_off += 1
            assert abs(got - ref) <= np.spacing(ref)
            exact = Decimal(x * y).ln()
            assert abs(Decimal(got) - exact) <= abs(Decimal(ref) - exact), (x, y)
    assert n_off < 10


def test_mi_from_table_matches_sklearn_golden(math_host, lntab, golden):
    n = 0
    for case in golden("tables.json"):
        t = u32(case["table"])
        if int(t.sum()) >= (1 << 17):
            continue
        got = math_host.t_mi_from_table(t.ctypes.data, lntab.ctypes.data)
        assert got == unhex(case["mi"]), case
        n += 1
    assert n > 2000


def test_mi_2x2_fast_path_equals_general(math_host, lntab):
    """The bi-allelic fast path and the 3x3 path give the same bits; both agree
    with the oracle (exactly, except where libm's log is not the rounded one)."""
    rng = np.random.default_rng(4)
    got, want = [], []
    for _ in range(5000):
        c = rng.integers(0, int(rng.choice([3, 12, 200, 5000])), 4)
        if c.sum() == 0:
            continue
        t = np.zeros(9, dtype=np.uint32)
        t[[4, 5, 7, 8]] = c
        a = math_host.t_mi_from_table(t.ctypes.data, lntab.ctypes.data)
        b = math_host.t_mi_from_2x2(int(c[0]), int(c[1]), int(c[2]), int(c[3]), lntab.ctypes.data)
        assert a == b
        got.append(a)
        want.append(c_oracle.mi_from_table(t.tolist()))
    assert_mi_close(got, want, min_exact=0.99)


def test_mi_from_table_random_vs_oracle(math_host, lntab):
    rng = np.random.default_rng(5)
    got, want = [], []
    for _ in range(20000):
        t = rng.integers(0, int(rng.choice([2, 5, 30, 300, 9000])), 9)
        t[rng.random(9) < rng.choice([0.0, 0.4, 0.7])] = 0
        if t.sum() == 0:
            continue
        got.append(math_host.t_mi_from_table(u32(t).ctypes.data, lntab.ctypes.data))
        want.append(c_oracle.mi_from_table(t.tolist()))
    assert_mi_close(got, want, min_exact=0.99)


def test_neumaier_mean_is_python_sum_over_len(math_host):
    rng = np.random.default_rng(6)
    for n in (1, 2, 3, 9, 49, 200):
        for _ in range(40):
            v = rng.random(n) * 10.0 ** rng.integers(-9, 3, n)
            want = sum(v.tolist()) / n
            assert math_host.t_neumaier_mean(v.ctypes.data, n) == want
    assert math.isnan(math_host.t_neumaier_mean(np.zeros(1).ctypes.data, 0))


def test_pair_index_roundtrip(math_host):
    import ctypes
    i, j = ctypes.c_uint32(), ctypes.c_uint32()
    for S in (2, 3, 4, 50, 51, 64, 65, 1000, 2000, 65535):
        npairs = S * (S - 1) // 2
        probe = set([0, npairs - 1] + np.random.default_rng(S).integers(0, npairs, 300).tolist())
        for k in range(S - 1):                       # every row boundary
            off = math_host.t_row_off(k, S)
            probe.update([off, max(0, off - 1)])
        for p in probe:
            math_host.t_pair_ij(int(p), S, ctypes.byref(i), ctypes.byref(j))
            assert 0 <= i.value < j.value < S
            assert math_host.t_row_off(i.value, S) + (j.value - i.value - 1) == p
    # combinations order
    S = 7
    want = [(a, b) for a in range(S) for b in range(a + 1, S)]
    got = []
    for p in range(len(want)):
        math_host.t_pair_ij(p, S, ctypes.byref(i), ctypes.byref(j))
        got.append((i.value, j.value))
    assert got == want
    # stepping through the triangle without the square root (k_pairs_generic / k_count: stride 256)
    for S in (2, 3, 24, 65, 700):
        pairs = [(a, b) for a in range(S) for b in range(a + 1, S)]
        for start in (0, 1, 7, len(pairs) // 2):
            for step in (1, 32, 256):
                if start >= len(pairs):
                    continue
                i.value, j.value = pairs[start]
                p = start
                while p + step < len(pairs):
                    math_host.t_pair_advance(ctypes.byref(i), ctypes.byref(j), S, step)
                    p += step
                    assert (i.value, j.value) == pairs[p]


def test_ecdf_y_is_numpy_linspace(math_host):
    for n in (1, 2, 3, 7, 10, 49, 1000, 99991):
        y = np.concatenate([[0.0], np.linspace(1 / n, 1, n)])
        idx = sorted(set([0, 1, n // 2, n - 1, n] + np.random.default_rng(n).integers(0, n + 1, 50).tolist()))
        for k in idx:
            assert math_host.t_ecdf_y(k, n) == y[k], (k, n)
    xs, y = oracle.ecdf_table([0.1, 0.2, 0.2, 0.4])
    assert [math_host.t_ecdf_y(k, 4) for k in range(5)] == y.tolist()
