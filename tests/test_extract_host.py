"""The native extraction (l-giremi_b200/extract.py: C++ cs scanner + interned reads + the site filters)
against the UNMODIFIED reference's get_region_mismatches_with_filters (baseline/_ref) on simulated
spliced reads: same surviving sites, same fields, same removed sites with the same reasons, same dict
orders -- with default filters and with filters tightened until every branch fires (window filter,
shallow alleles, homopolymers, repeats).  Host only: no GPU needed."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import simdata  # noqa: E402

extract = importlib.import_module("l-giremi_b200.extract")
enc = importlib.import_module("l-giremi_b200.encode")


@pytest.fixture(scope="module")
def dataset():
    return simdata.Dataset(seed=20261041, n_genes=5, reads_per_gene=140, noise=0.004)


def same_extraction(ref_giremi, ds, fp, **kw):
    chrom, a, b, _n = fp
    snps = [p for p in ds.snp_positions if a <= p < b]
    want, want_removed = ref_giremi.mismatch.get_region_mismatches_with_filters(
        chromosome=chrom, start_pos=a, end_pos=b, sam=ds.sam(), genome=ds.fasta(), snp_positions=snps,
        **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
    got, got_removed, names = extract.get_region_sites(
        chrom, a, b, ds.sam(), ds.fasta(), snp_positions=snps,
        **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
    n_sites = 0
    for s in '+-':
        assert list(got[s]) == list(want[s]), "site order on %s" % s
        for pos in want[s]:
            w, g = want[s][pos], got[s][pos]
            assert (g['ref'], g['type'], g['up'], g['down']) == (w['ref'], w['type'], w['up'], w['down'])
            assert list(g['depth'].items()) == list(w['depth'].items())
            assert list(g['neighbor'].items()) == list(w['neighbor'].items())
            assert list(g['nt']) == list(w['nt'])
            for nt in w['nt']:
                assert [names[k] for k in g['nt'][nt]] == w['nt'][nt]
            n_sites += 1
        assert list(got_removed[s]) == list(want_removed[s]), "removed order on %s" % s
        for pos in want_removed[s]:
            assert got_removed[s][pos]['removed'] == want_removed[s][pos]['removed']
            assert list(got_removed[s][pos]['depth'].items()) == list(want_removed[s][pos]['depth'].items())
    reasons = {}
    for s in '+-':
        for site in want_removed[s].values():
            reasons[site['removed']] = reasons.get(site['removed'], 0) + 1
    return n_sites, reasons, got


def test_native_extraction_equals_reference_default_filters(ref_giremi, dataset):
    total = 0
    for fp in dataset.footprints(2):
        n, reasons, _ = same_extraction(ref_giremi, dataset, fp, min_total_depth=2)
        total += n
    assert total > 30


def test_native_extraction_equals_reference_every_filter_fires(ref_giremi, dataset):
    seen = {}
    fps = dataset.footprints(2)
    for fp in fps:
        chrom, a, b, _n = fp
        repeats = [[a + 200, a + 900], [a + 850, a + 1300], [b - 400, b - 100]]      # overlapping: the helper sorts, no merge
        n, reasons, _ = same_extraction(
            ref_giremi, dataset, fp, min_total_depth=4, min_allele_depth=2, min_allele_ratio=0.02,
            mismatch_window_size=60, max_window_mismatch=1, max_window_mismatch_type=1, homopoly_length=2,
            min_dist_from_splice=7, simple_repeat_intervals=repeats, min_het_snp_ratio=0.4, max_het_snp_ratio=0.6)
        for k, v in reasons.items():
            seen[k] = seen.get(k, 0) + v
    for reason in ('too many window mismatches', 'too few usable reads after filters', 'not enough allele after filters',
                   'in homopoly regions', 'in simple repeat regions'):
        assert seen.get(reason, 0) > 0, (reason, seen)
    # non-spliced reads kept, no splice filter
    same_extraction(ref_giremi, dataset, fps[0], keep_non_spliced_read=True, min_dist_from_splice=0, min_total_depth=2)


def test_indexed_encoder_equals_the_dict_encoder(ref_giremi, dataset):
    """Planes written from read indices == planes of the reference's dict through the label-matrix encoder, up to
    the order of the reads (compared through per-site allele counts and pairwise co-occurrence counts)."""
    checked = 0
    for fp in dataset.footprints(2):
        chrom, a, b, _n = fp
        snps = [p for p in dataset.snp_positions if a <= p < b]
        want, _ = ref_giremi.mismatch.get_region_mismatches_with_filters(
            chromosome=chrom, start_pos=a, end_pos=b, sam=dataset.sam(), genome=dataset.fasta(), snp_positions=snps,
            min_total_depth=2)
        got, _, _names = extract.get_region_sites(chrom, a, b, dataset.sam(), dataset.fasta(), snp_positions=snps,
                                                 min_total_depth=2)
        for s in '+-':
            if len(want[s]) < 2:
                continue
            pb = extract.encode_indexed(got[s])
            eu = enc.encode_mismatches(want[s])
            ref_pb = enc.pack_units([eu])
            assert pb.positions[0] == eu.positions and pb.types[0] == eu.types
            assert np.array_equal(pb.site_flags, ref_pb.site_flags)
            S, W = eu.n_sites, int(pb.units['row_words'][0])
            assert int(pb.units['n_reads'][0]) == eu.n_reads

            def label_matrix(batch):
                w = int(batch.units['row_words'][0])
                bits = np.unpackbits(batch.planes.reshape(S, 3, w).view(np.uint8), axis=-1, bitorder='little')
                return bits[:, 0].astype(np.int32) * 2 + bits[:, 1] + (bits[:, 2] - bits[:, 0] - bits[:, 1]) * 4
            la, lb = label_matrix(pb), label_matrix(ref_pb)      # 2 major, 1 minor, 4 other, 0 uncovered
            # the same multiset of read columns
            ca = sorted(map(bytes, la[:, :eu.n_reads].T.astype(np.uint8)))
            cb = sorted(map(bytes, lb[:, :eu.n_reads].T.astype(np.uint8)))
            assert ca == cb
            checked += 1
    assert checked >= 4
