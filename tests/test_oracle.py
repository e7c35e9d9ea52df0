"""Pins the CPU oracle (oracle/oracle.py, oracle/oracle_mi.c) against the golden
fixtures that tests/golden/make_golden.py produced by executing the real
reference (gxiaolab/L-GIREMI v0.2.4) and the installed scikit-learn.

Bit-exact: floats in the fixtures are C99 hex strings."""
import math
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_mismatches, unhex

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import c_oracle  # noqa: E402
import oracle  # noqa: E402


def same_bits(a, b):
    return (math.isnan(a) and math.isnan(b)) or a == b


# --------------------------------------------------------------------------- tables
def test_mi_from_table_python_matches_sklearn(golden):
    for case in golden("tables.json"):
        t = np.array(case["table"]).reshape(3, 3)
        assert oracle.mi_from_table(t) == unhex(case["mi"]), case


def test_mi_from_table_c_matches_sklearn(golden):
    for case in golden("tables.json"):
        assert c_oracle.mi_from_table(case["table"]) == unhex(case["mi"]), case


# --------------------------------------------------------------------------- KATs
def test_known_answers(golden):
    kat = golden("kat.json")
    expect = {"ab": 0.6931471805599454, "cb": 0.4620981203732968, "db": 0.2157615543388357}
    for name, value in expect.items():
        rows = kat[name]["rows"]
        assert len(rows) == 1 and unhex(rows[0][4]) == value
    assert len(kat["ef_min6"]["rows"]) == 1 and unhex(kat["ef_min6"]["rows"][0][4]) == 0.0
    assert kat["ef_min7"]["rows"] == []
    assert kat["index_error"] is True
    means = {p: unhex(m) for p, m in kat["abc"]["means"]}
    assert means == {10: 0.6931471805599454, 20: 0.5776226504666211, 30: 0.4620981203732968}


@pytest.mark.parametrize("impl", ["python", "c"])
def test_kat_units(golden, impl):
    kat = golden("kat.json")
    for name in ("ab", "cb", "db", "ef_min6", "ef_min7", "abc"):
        check_unit(kat[name], impl)


def check_unit(unit, impl):
    m = golden_mismatches(unit)
    mc = unit["min_common"]
    want_rows = [[r[0], r[1], r[2], r[3], unhex(r[4])] for r in unit["rows"]]
    want_kept = [[r[0], r[1], r[2], r[3], unhex(r[4])] for r in unit["kept_rows"]]
    want_means = [[p, unhex(v)] for p, v in unit["means"]]
    if impl == "python":
        full, kept, means = oracle.mi_step(m, mc)
        assert full == want_rows
        assert kept == want_kept
        assert means == want_means
    else:
        out = c_oracle.unit_step(m, mc)
        assert out["rows"] == want_rows
        got = {p: v for p, v, c in zip(out["positions"], out["mean"].tolist(), out["cnt"].tolist()) if c}
        assert got == dict((p, v) for p, v in want_means)
        # sites outside every kept pair have NaN (mismatch.py:476-479)
        for p, v, c in zip(out["positions"], out["mean"].tolist(), out["cnt"].tolist()):
            assert (c == 0) == math.isnan(v)


@pytest.mark.parametrize("impl", ["python", "c"])
@pytest.mark.parametrize("fixture", ["units_fuzz.json", "units_synth.json"])
def test_units_match_reference(golden, impl, fixture):
    units = golden(fixture)
    assert len(units) >= 7
    for unit in units:
        check_unit(unit, impl)


def test_index_error_like_reference():
    site_ok = {'ref': 'A', 'type': 'het_snp', 'depth': {'A': 6, 'C': 6},
               'nt': {'A': ['r%d' % k for k in range(6)], 'C': ['r%d' % k for k in range(6, 12)]}}
    site_bad = {'ref': 'A', 'type': 'mismatch', 'depth': {'G': 12},
                'nt': {'G': ['r%d' % k for k in range(12)]}}
    with pytest.raises(IndexError):
        oracle.pair_mutual_info({10: site_bad, 20: site_ok}, 6)
    with pytest.raises(IndexError):
        c_oracle.unit_step({10: site_bad, 20: site_ok}, 6)
    # the pair is dropped before the ranking is looked at -> no error
    assert oracle.pair_mutual_info({10: site_bad, 20: site_ok}, 13) == []
    assert c_oracle.unit_step({10: site_bad, 20: site_ok}, 13)["rows"] == []


# --------------------------------------------------------------------------- sums
def test_python_sum_matches_builtin():
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 3, 10, 57):
        v = (rng.random(n) * 10.0 ** rng.integers(-8, 8, n)).tolist()
        assert oracle.python_sum_order(v) == sum(v)
        if n:
            assert c_oracle.python_sum(v) == sum(v)


def test_numpy_sum_order_matches_numpy():
    rng = np.random.default_rng(2)
    for n in range(0, 10):
        for _ in range(50):
            v = rng.random(n) * 10.0 ** rng.integers(-6, 6, n)
            assert oracle.numpy_sum_order(v.tolist()) == float(v.sum())


# --------------------------------------------------------------------------- ecdf / mip / calls
def test_ecdf_functions(golden):
    for case in golden("ecdf.json")["functions"]:
        x = np.array([unhex(v) for v in case["x"]])
        xs, y = oracle.ecdf_table(x)
        for s, want in zip(case["samples"], case["y"]):
            got = y[np.searchsorted(xs, unhex(s), side="left")]
            assert got == unhex(want)


def test_ecdf_kat(golden):
    k = golden("kat.json")["ecdf"]
    xs, y = oracle.ecdf_table(k["x"])
    got = [float(y[np.searchsorted(xs, s)]) for s in k["samples"]]
    assert got == [unhex(v) for v in k["y"]] == [0, 0, 0.25, 0.75, 0.75, 1.0]


@pytest.mark.parametrize("impl", ["python", "c"])
def test_site_table_mip_and_calls(golden, impl):
    tab = golden("ecdf.json")["site_table"]
    mean = np.array([unhex(v) for v in tab["mean"]])
    types = tab["type"]
    want_mip = np.array([unhex(v) for v in tab["mip"]])
    code = np.array([c_oracle.TYPE_CODE[t] for t in types], dtype=np.uint8)
    if impl == "python":
        mip = oracle.mip_values(mean, code == 2)
        call = oracle.threshold_calls(mean, mip, code == 0, tab["threshold"])
    else:
        mip, call = c_oracle.mip_calls(mean, code, tab["threshold"])
    assert all(same_bits(a, b) for a, b in zip(mip.tolist(), want_mip.tolist()))
    assert (call == 1).tolist() == tab["positive"]
    assert (call == 2).tolist() == tab["negative"]


# --------------------------------------------------------------------------- timing port
def test_timing_port_matches_oracle(golden):
    import ref_port
    for unit in golden("units_fuzz.json")[:30]:
        m = golden_mismatches(unit)
        rows = ref_port.port_pair_mutual_info(m, unit["min_common"])
        assert rows == [[r[0], r[1], r[2], r[3], unhex(r[4])] for r in unit["rows"]]
