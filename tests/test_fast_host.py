"""The small-unit ("fast") path's device functions (csrc/lgmi_fast.cuh), built
for the host: they must give the SAME BITS as the straightforward arithmetic in
csrc/lgmi_math.cuh (which test_math_host.py pins to sklearn's golden values)."""
import numpy as np

from conftest import unhex

NTAB = 257          # the kernel's shared-memory table covers counts up to 256 reads


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def test_markstein_quotient_is_the_rounded_quotient(fast_host):
    """q0 = n*RN(1/N); r = fma(-q0,N,n); q = fma(r,RN(1/N),q0) == RN(n/N) for all 0<=n<=N<=3000."""
    assert fast_host.f_markstein_mismatches(3000) == 0


def test_u32_to_double_trick(fast_host):
    import ctypes
    fast_host.f_u32_to_double.restype = ctypes.c_double
    fast_host.f_u32_to_double.argtypes = [ctypes.c_uint32]
    for n in list(range(0, 600)) + [65535, 65536, 2**31 - 1, 2**31, 2**32 - 1]:
        assert fast_host.f_u32_to_double(n) == float(n)


def test_carry_save_popcount(fast_host):
    rng = np.random.default_rng(1)
    for nw in (2, 4, 7, 8):
        for _ in range(3000):
            x = rng.integers(0, 1 << 32, 8, dtype=np.uint64).astype(np.uint32)
            y = rng.integers(0, 1 << 32, 8, dtype=np.uint64).astype(np.uint32)
            if rng.random() < 0.2:
                x[:] = 0xffffffff
                y[:] = 0xffffffff
            want = sum(bin(int(a) & int(b)).count("1") for a, b in zip(x[:nw], y[:nw]))
            assert fast_host.f_and_popc(nw, x.ctypes.data, y.ctypes.data) == want


def test_pair_counts_packing(fast_host):
    rng = np.random.default_rng(2)
    for nw in (2, 4, 7, 8):
        for _ in range(500):
            ri = np.zeros(20, np.uint32)
            rj = np.zeros(20, np.uint32)
            for row in (ri, rj):
                P = rng.integers(0, 1 << 32, 8, dtype=np.uint64).astype(np.uint32)
                M = P & rng.integers(0, 1 << 32, 8, dtype=np.uint64).astype(np.uint32)
                P[nw:] = 0
                M[nw:] = 0
                row[:8], row[8:16] = M, P
            v = fast_host.f_pair_counts(nw, ri.ctypes.data, rj.ctypes.data, 0, 0)
            pc = lambda a, b: int(sum(bin(int(x) & int(y)).count("1") for x, y in zip(a, b)))
            want = (pc(ri[8:16], rj[8:16]), pc(ri[:8], rj[8:16]), pc(ri[8:16], rj[:8]), pc(ri[:8], rj[:8]))
            got = (v & 511, (v >> 9) & 511, (v >> 18) & 511, (v >> 27) & 511)
            assert v >> 36 == 0
            assert got == want
            # the min-common early exit: |Pi&Pj| + "other" reads against the threshold, strict '<'
            none = (1 << 64) - 1
            for n_other in (0, 3):
                n = want[0] + n_other
                assert fast_host.f_pair_counts(nw, ri.ctypes.data, rj.ctypes.data, n_other, n) == v
                assert fast_host.f_pair_counts(nw, ri.ctypes.data, rj.ctypes.data, n_other, n + 1) == none


def test_fast_2x2_same_bits_as_reference_arithmetic(fast_host, math_host, lntab):
    rng = np.random.default_rng(3)
    cells = []
    for _ in range(40000):
        c = rng.integers(0, int(rng.choice([2, 4, 12, 60, 120])), 4)
        c[rng.random(4) < rng.choice([0.0, 0.3])] = 0
        if 0 < c.sum() < NTAB:
            cells.append(c)
    cells = u32(np.array(cells))
    out = np.empty(len(cells))
    fast_host.f_mi_2x2_many(cells.ctypes.data, len(cells), lntab.ctypes.data, NTAB, out.ctypes.data)
    for c, got in zip(cells.tolist(), out.tolist()):
        want = math_host.t_mi_from_2x2(c[0], c[1], c[2], c[3], lntab.ctypes.data)
        assert got == want, c


def test_fast_3x3_same_bits_as_reference_arithmetic(fast_host, math_host, lntab):
    rng = np.random.default_rng(4)
    tabs = []
    for _ in range(40000):
        t = rng.integers(0, int(rng.choice([2, 3, 10, 40, 56])), 9)
        if rng.random() < 0.6:                       # "other" cells are small in practice
            t[[0, 1, 2, 3, 6]] = rng.integers(0, 3, 5)
        t[rng.random(9) < rng.choice([0.0, 0.3, 0.6])] = 0
        if 0 < t.sum() < NTAB:
            tabs.append(t)
    tabs = u32(np.array(tabs))
    out = np.empty(len(tabs))
    fast_host.f_mi_3x3_many(tabs.ctypes.data, len(tabs), lntab.ctypes.data, NTAB, out.ctypes.data)
    n8 = 0
    for t, got in zip(tabs, out.tolist()):
        want = math_host.t_mi_from_table(t.ctypes.data, lntab.ctypes.data)
        assert got == want, t
        n8 += int((t != 0).sum() >= 8)
    assert n8 > 500                                  # numpy's pairwise branch was exercised


def test_fast_paths_match_sklearn_golden(fast_host, lntab, golden):
    n = 0
    for case in golden("tables.json"):
        t = u32(case["table"])
        if int(t.sum()) >= NTAB:
            continue
        assert fast_host.f_mi_3x3(t.ctypes.data, lntab.ctypes.data, NTAB) == unhex(case["mi"]), case
        if not t[[0, 1, 2, 3, 6]].any():
            assert fast_host.f_mi_2x2(int(t[4]), int(t[5]), int(t[7]), int(t[8]), lntab.ctypes.data, NTAB) == \
                unhex(case["mi"]), case
        n += 1
    assert n > 1000


def test_global_table_epilogues_same_bits_for_large_counts(fast_host, math_host, lntab, golden):
    """k_tile_finish runs the same branch-free epilogues over the context's full ln table and the correctly
    rounded reciprocal (GlobalTab): counts far beyond the 256 of the shared-memory table must still give the
    bits of the straightforward arithmetic, and the golden sklearn tables of any size must reproduce."""
    big = lntab                                       # 2^17 entries
    rng = np.random.default_rng(9)
    tabs = []
    for _ in range(30000):
        scale = int(rng.choice([3, 40, 300, 2000, 7000]))
        t = rng.integers(0, scale, 9)
        if rng.random() < 0.6:
            t[[0, 1, 2, 3, 6]] = rng.integers(0, 4, 5)
        t[rng.random(9) < rng.choice([0.0, 0.3, 0.6])] = 0
        if 0 < t.sum() < 65536:
            tabs.append(t)
    tabs.append(np.array([0, 0, 0, 0, 65535, 0, 0, 0, 0]))
    tabs.append(np.array([0, 0, 0, 0, 30000, 2, 0, 1, 35000]))
    tabs = u32(np.array(tabs))
    out = np.empty(len(tabs))
    fast_host.g_mi_3x3_many(tabs.ctypes.data, len(tabs), big.ctypes.data, out.ctypes.data)
    for t, got in zip(tabs, out.tolist()):
        assert got == math_host.t_mi_from_table(t.ctypes.data, big.ctypes.data), t
    two = tabs[~tabs[:, [0, 1, 2, 3, 6]].any(axis=1)]
    cells = u32(two[:, [4, 5, 7, 8]])
    out2 = np.empty(len(cells))
    fast_host.g_mi_2x2_many(cells.ctypes.data, len(cells), big.ctypes.data, out2.ctypes.data)
    for c, got in zip(cells.tolist(), out2.tolist()):
        assert got == math_host.t_mi_from_2x2(c[0], c[1], c[2], c[3], big.ctypes.data), c
    assert len(cells) > 100
    n = 0
    for case in golden("tables.json"):
        t = u32(case["table"])
        if int(t.sum()) >= (1 << 17):
            continue
        one = np.empty(1)
        fast_host.g_mi_3x3_many(t.ctypes.data, 1, big.ctypes.data, one.ctypes.data)
        assert one[0] == unhex(case["mi"]), case
        n += 1
    assert n > 2000


def test_global_table_epilogues_same_bits_up_to_two_million_reads(fast_host, math_host):
    """The deep-unit pair kernel (k_pairs_generic<2>) runs the reorganised epilogues for units of up to
    kFastMathMaxCount = 2^21 - 1 reads (cfg3: 100 000): same bits as the straightforward arithmetic on tables with
    counts of that size -- the Markstein quotient is the rounded quotient for any integers below 2^32, the logs
    come from the same double-double table."""
    n = 1 << 21
    big = np.empty(2 * n, dtype=np.float64)
    math_host.t_build_lntab(big.ctypes.data, 0, n)
    rng = np.random.default_rng(10)
    tabs = []
    for _ in range(20000):
        scale = int(rng.choice([20000, 100000, 230000]))
        t = rng.integers(0, scale, 9)
        if rng.random() < 0.7:                                       # "other" cells stay small
            t[[0, 1, 2, 3, 6]] = rng.integers(0, int(rng.choice([2, 30, 800])), 5)
        t[rng.random(9) < rng.choice([0.0, 0.3, 0.6])] = 0
        if 0 < t.sum() < n:
            tabs.append(t)
    tabs.append(np.array([0, 0, 0, 0, n - 1, 0, 0, 0, 0]))
    tabs.append(np.array([1, 0, 3, 0, 700000, 2, 5, 1, 1300000]))
    tabs = u32(np.array(tabs))
    out = np.empty(len(tabs))
    fast_host.g_mi_3x3_many(tabs.ctypes.data, len(tabs), big.ctypes.data, out.ctypes.data)
    for t, got in zip(tabs, out.tolist()):
        assert got == math_host.t_mi_from_table(t.ctypes.data, big.ctypes.data), t
    two = tabs[~tabs[:, [0, 1, 2, 3, 6]].any(axis=1)]
    cells = u32(two[:, [4, 5, 7, 8]])
    out2 = np.empty(len(cells))
    fast_host.g_mi_2x2_many(cells.ctypes.data, len(cells), big.ctypes.data, out2.ctypes.data)
    for c, got in zip(cells.tolist(), out2.tolist()):
        assert got == math_host.t_mi_from_2x2(c[0], c[1], c[2], c[3], big.ctypes.data), c
    assert len(cells) > 100 and len(tabs) > 15000
    assert fast_host.f_markstein_large_mismatches(200000, 11) == 0


def test_two_sum_error_equals_cpython_compensation_term():
    """csrc/lgmi_fast_kernel.cuh two_sum_err: the branch-free six-addition error of RN(s + x) is, bit for bit, the term
    CPython's compensated float sum adds -- (s - t) + x when |s| >= |x|, (x - t) + s otherwise (Python/bltinmodule.c
    cs_add, what mutual_information.py:56-58 runs) -- over magnitudes from denormal-ish MI noise to large sums."""
    rng = np.random.default_rng(12)
    n = 400000
    s = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-18, 6, n)
    x = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-18, 6, n)
    s[::7] = 0.0
    x[::11] = 0.0
    x[::13] = s[::13]
    t = s + x
    big, small = np.where(s >= x, s, x), np.where(s >= x, x, s)
    want = (big - t) + small
    bp = t - s
    got = (s - (t - bp)) + (x - bp)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # and a whole sum: same (s, c) sequence as math.fsum-free CPython sum() of non-negative values
    vals = (np.abs(rng.standard_normal(3000)) * 10.0 ** rng.integers(-12, 1, 3000)).tolist()
    acc_s = acc_c = 0.0
    for v in vals:
        tt = acc_s + v
        b = tt - acc_s
        acc_c += (acc_s - (tt - b)) + (v - b)
        acc_s = tt
    total = acc_s + acc_c if acc_c != 0.0 else acc_s
    assert total == sum(vals)
