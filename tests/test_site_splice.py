"""Site x splice-site MI: the reference's second mutual_info_score call
site, giremi/script/calculate_site_splice_mi.py:106-125."""
import importlib
import os
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT, assert_mi_close, unhex

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402


def random_case(rng, n_sites=30, n_splices=12):
    sites, splices, pairs = {}, {}, []
    reads = ["read%05d" % k for k in range(4000)]
    for s in range(n_splices):
        splices["chr1:%d" % (5000 + 97 * s)] = list(rng.choice(reads, int(rng.integers(1, 900)), replace=False))
    for k in range(n_sites):
        label = "chr1:%d" % (1000 + 37 * k)
        n = int(rng.integers(1, 700))
        cover = list(rng.choice(reads, n, replace=False))
        alleles = {}
        for name in cover:
            alleles.setdefault(str(rng.choice(list("ACGT"), p=[0.5, 0.3, 0.15, 0.05])), []).append(name)
        if n > 3 and rng.random() < 0.3:                       # a read listed twice, under two alleles
            alleles.setdefault('A', []).append(cover[0])
            alleles.setdefault('G', []).append(cover[0])
        sites[label] = alleles
        for seq in alleles:
            for sp in rng.choice(list(splices), int(rng.integers(1, 5)), replace=False):
                pairs.append((label, seq, str(sp)))
    pairs.append((label, seq, "chr1:999999"))                 # a splice label nobody has: one class -> 0.0
    return sites, splices, pairs


def golden_tables(golden, tmp_path):
    """The two TSVs the golden vectors were generated from (tests/golden/make_site_splice_golden.py)."""
    g = golden("site_splice.json")
    sf, pf = str(tmp_path / "site.tsv"), str(tmp_path / "splice.tsv")
    pd.DataFrame(g["site_rows"], columns=["read_name", "chromosome", "pos", "seq"]).to_csv(sf, sep="\t", index=False)
    pd.DataFrame(g["splice_rows"], columns=["read_name", "chromosome", "pos", "type", "corrected_pos", "annotation"]
                 ).to_csv(pf, sep="\t", index=False)
    return g, sf, pf


def test_oracle_and_table_builder_against_the_reference_script(golden, tmp_path):
    """PIN: tests/golden/site_splice.json holds what the reference's own calculate_site_splice_mi.py main()
    wrote for these two tables.  The host-side table builder must list the same pairs with the same counts in
    the same order, and the oracle must reproduce every MI bit for bit."""
    ss = importlib.import_module("l-giremi_b200.site_splice")
    g, sf, pf = golden_tables(golden, tmp_path)
    rows, sites, splices = ss.read_tables(sf, pf)
    want = g["pairs"]
    assert [[r[0], int(r[1]), r[2], int(r[3]), r[4]] for r in rows] == [p[:5] for p in want]
    triples = [(r[0] + ':' + r[1], r[2], r[0] + ':' + r[3]) for r in rows]
    got = oracle.site_splice_mutual_info(sites, splices, triples)
    assert got == [unhex(p[5]) for p in want]
    assert max(got) == pytest.approx(np.log(2), abs=1e-3) and min(got) == 0.0     # an exact link and a one-class pair


@pytest.mark.gpu
def test_site_splice_cli_drop_in_against_the_reference_script(lg, gpu_ctx, golden, tmp_path):
    """The drop-in for the console script: same output table as the reference wrote."""
    ss = importlib.import_module("l-giremi_b200.site_splice")
    g, sf, pf = golden_tables(golden, tmp_path)
    prefix = str(tmp_path / "out")
    ss.main(["-m", sf, "-s", pf, "-o", prefix])
    table = pd.read_table(prefix + ".site_splice_pair", dtype={"chromosome": str, "seq": str}, float_precision="round_trip")
    assert list(table.columns) == g["pair_columns"]
    want = g["pairs"]
    assert [[str(r.chromosome), int(r.site_pos), str(r.seq), int(r.splice_pos), int(r["count"])]
            for _, r in table.iterrows()] == [p[:5] for p in want]
    assert_mi_close(table["mi"].to_numpy(), [unhex(p[5]) for p in want], 0.9, "site x splice MI")


def test_oracle_follows_the_script_literally():
    """The restatement against the script's own expressions (with_seq / with_splice lists
    handed to sklearn), on a small case."""
    from sklearn.metrics import mutual_info_score
    rng = np.random.default_rng(4)
    sites, splices, pairs = random_case(rng, 6, 4)
    got = oracle.site_splice_mutual_info(sites, splices, pairs)
    for (site_label, seq, splice_label), mi in zip(pairs, got):
        read_site_seq = sites[site_label][seq]
        read_splice = splices.get(splice_label, [])
        read_site_all = sorted(sum((sites[site_label][c] for c in sites[site_label]), []))
        with_seq = [int(a in read_site_seq) for a in read_site_all]
        with_splice = [int(a in read_splice) for a in read_site_all]
        assert mi == mutual_info_score(with_seq, with_splice)


@pytest.mark.gpu
def test_site_splice_mi_one_submit(lg, gpu_ctx):
    rng = np.random.default_rng(5)
    sites, splices, pairs = random_case(rng)
    want = oracle.site_splice_mutual_info(sites, splices, pairs)
    before = gpu_ctx.launch_count
    got = lg.site_splice_mutual_info(sites, splices, pairs, ctx=gpu_ctx)
    assert gpu_ctx.launch_count - before <= 12
    assert len(got) == len(pairs) > 100
    assert_mi_close(got, want, 0.95, "site x splice MI")
    assert got[-1] == 0.0
    assert lg.site_splice_mutual_info(sites, splices, [], ctx=gpu_ctx) == []
