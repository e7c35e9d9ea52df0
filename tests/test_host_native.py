"""Native host pieces of SURVEY 8f (f1 encoder, f2 cs-tag scanner): no device needed."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_mismatches

sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fuzz import random_mismatches  # noqa: E402

enc = importlib.import_module("l-giremi_b200.encode")


def test_cs_scan_matches_reference_golden(lg, golden):
    for case in golden("cs_scan.json"):
        mm, introns = enc.cs_read_mismatches(case["cs"], case["start"], case["min_dist"])
        assert mm == case["mismatches"], case["cs"]
        assert introns == case["introns"], case["cs"]


def test_cs_scan_rejects_what_the_reference_rejects(lg):
    for bad in ("=ACGT", ":12x", "*a", ":", "~gtag", "12", "!3"):
        with pytest.raises(ValueError):
            enc.cs_read_mismatches(bad, 0, 4)
    assert enc.cs_read_mismatches("", 5, 4) == ([], [])


def test_cs_scan_on_simulated_reads_against_installed_reference(lg, ref_giremi):
    """Every read of a simulated dataset: the reference's CS class + its FILTER 1 (run here from
    baseline/_ref) against the native scanner."""
    import simdata
    from giremi.cs import CS
    from giremi.utils import merge_intervals, positions_in_intervals
    ds = simdata.Dataset(seed=5, n_genes=3, reads_per_gene=60)
    n = 0
    for r in ds.reads:
        c = CS.from_cs_tag_string(r.cs, "chr1", r.reference_start, "+")
        mm = sorted([[a[0], a[3]] for a in c.get_mismatches(coordinate='contig')], key=lambda a: a[0])
        intr = sorted(c.get_introns(coordinate='contig'), key=lambda a: a[0])
        if mm and intr:
            pos = sorted([a[0] for a in intr] + [a[1] for a in intr])
            iv, _ = merge_intervals([[a - 4, a + 4] for a in pos])
            inside, _ = positions_in_intervals([a[0] for a in mm], iv)
        else:
            inside = [False] * len(mm)
        want = [[p, v.upper()] for (p, v), i in zip(mm, inside) if not i]
        got, got_intr = enc.cs_read_mismatches(r.cs, r.reference_start, 4)
        assert got == want and got_intr == [[a[0], a[1]] for a in intr]
        n += len(want)
    assert n > 100


def test_native_encoder_equals_python_encoder(lg, golden):
    """lgmi_encode_unit on the golden and fuzzed dicts: the same planes, flags, read count and
    bad-site set as encode_mismatches + pack_units (duplicate read names, depth ties, third
    alleles, depth != list length)."""
    rng = np.random.default_rng(8)
    dicts = [golden_mismatches(u) for u in golden("units_fuzz.json") + golden("units_synth.json")]
    dicts += [random_mismatches(rng) for _ in range(200)]
    dicts += [{}, {7: {'ref': 'A', 'type': 'snp', 'depth': {'A': 2}, 'nt': {'A': ['x', 'y']}}}]
    for m in dicts:
        eu = enc.encode_mismatches(m)
        want = enc.pack_units([eu])
        got = enc.encode_mismatches_native(m)
        assert got.units.tolist() == want.units.tolist()
        assert np.array_equal(got.planes, want.planes)
        assert np.array_equal(got.site_flags, want.site_flags)
        assert got.bad_sites[0] == eu.bad_sites and got.positions[0] == eu.positions and got.types[0] == eu.types


synth = importlib.import_module("l-giremi_b200.synth")


def test_packed_two_plane_form_round_trips(lg):
    """PlaneBatch.packed2(): 2 bits per (site, read) hold exactly the three planes."""
    pb, _ = synth.make_heavy_tail(20261027, 40, s_max=120, r_max=900)
    p2 = pb.packed2()
    assert p2.size * 3 == pb.planes.size * 2
    for k in range(pb.n_units):
        S, W = int(pb.units['n_sites'][k]), int(pb.units['row_words'][k])
        three = pb.planes[int(pb.units['plane_off'][k]):][:3 * S * W].reshape(S, 3, W)
        two = p2[int(pb.units['plane_off'][k]) // 3 * 2:][:2 * S * W].reshape(S, 2, W)
        assert np.array_equal(two[:, 0] & ~two[:, 1], three[:, 0])
        assert np.array_equal(two[:, 1] & ~two[:, 0], three[:, 1])
        assert np.array_equal(two[:, 0] | two[:, 1], three[:, 2])


def test_concat_plane_batches_equals_pack_units(lg):
    rng = np.random.default_rng(9)
    ms = [random_mismatches(rng) for _ in range(12)] + [{}]
    parts = [enc.encode_mismatches_native(m) for m in ms]
    got = enc.concat_plane_batches(parts)
    want = enc.pack_units([enc.encode_mismatches(m) for m in ms])
    assert got.units.tolist() == want.units.tolist()
    assert np.array_equal(got.planes, want.planes) and np.array_equal(got.site_flags, want.site_flags)
    assert got.positions == want.positions and got.types == want.types
