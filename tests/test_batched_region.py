"""The batched seam (SURVEY 8b ii) against the UNMODIFIED reference: the reference's
own region_mismatch_analysis (baseline/_ref, CPU) and l-giremi_b200.batched (GPU MI)
on the same simulated reads must give the same three DataFrames."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import MI_ATOL, MI_RTOL, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import simdata  # noqa: E402

pytestmark = pytest.mark.gpu

FILTERS = dict(min_total_depth=2, min_allele_ratio=0.05, min_allele_depth=3)


def frames_equal(got, want):
    for g, w, floats in zip(got, want, (("mean_mi",), ("mi",), ())):
        assert list(g.columns) == list(w.columns) and len(g) == len(w)
        exact = [c for c in w.columns if c not in floats]
        pd.testing.assert_frame_equal(g[exact].reset_index(drop=True), w[exact].reset_index(drop=True), check_exact=True)
        for c in floats:
            a, b = g[c].to_numpy(dtype=float), w[c].to_numpy(dtype=float)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            ok = np.isnan(b) | (np.abs(a - b) <= MI_RTOL * np.abs(b) + MI_ATOL)
            assert ok.all(), (c, a[~ok][:3], b[~ok][:3])


@pytest.fixture(scope="module")
def dataset():
    return simdata.Dataset(seed=20261025, n_genes=6, reads_per_gene=150)


def snps_in(ds, a, b):
    return [p for p in ds.snp_positions if a <= p < b]


def test_region_drop_in_and_one_submit_for_all_regions(lg, gpu_ctx, ref_giremi, dataset):
    ds = dataset
    fps = ds.footprints(2)
    assert len(fps) == 6
    want, regs = [], []
    n_pairs = 0
    for chrom, a, b, _n in fps:
        want.append(ref_giremi.mismatch.region_mismatch_analysis(
            chrom, a, b, ds.sam(), ds.fasta(), snp_positions=snps_in(ds, a, b), min_common_reads=6, **FILTERS))
        got = lg.region_mismatch_analysis(chrom, a, b, ds.sam(), ds.fasta(), snp_positions=snps_in(ds, a, b),
                                          min_common_reads=6, **FILTERS)
        frames_equal(got, want[-1])
        n_pairs += len(want[-1][1])
        regs.append(lg.extract_region(chrom, a, b, ds.sam(), ds.fasta(), snp_positions=snps_in(ds, a, b), **FILTERS))
    assert n_pairs > 20                                             # the comparison is not vacuous
    before = gpu_ctx.launch_count
    all_at_once = lg.analyse_extracted(regs, 6, ctx=gpu_ctx)
    assert 0 < gpu_ctx.launch_count - before <= 12                  # one submit, not one per region
    for got, w in zip(all_at_once, want):
        frames_equal(got, w)


def test_install_batched_rebinds_region_analysis(lg, gpu_ctx, ref_giremi, dataset):
    ds = dataset
    chrom, a, b, _n = ds.footprints(2)[2]
    kw = dict(snp_positions=snps_in(ds, a, b), min_common_reads=4, **FILTERS)
    stock = ref_giremi.mismatch.region_mismatch_analysis(chrom, a, b, ds.sam(), ds.fasta(), **kw)
    done = lg.install(batched=True)
    try:
        assert ("giremi.mismatch", "region_mismatch_analysis") in done
        assert ref_giremi.mismatch.region_mismatch_analysis is lg.region_mismatch_analysis
        patched = ref_giremi.mismatch.region_mismatch_analysis(chrom, a, b, ds.sam(), ds.fasta(), **kw)
    finally:
        lg.uninstall()
    assert ref_giremi.mismatch.region_mismatch_analysis is not lg.region_mismatch_analysis
    frames_equal(patched, stock)
    # the name-level patch alone (unbatched seam): the reference's own loop calling the GPU functions
    lg.install()
    try:
        seam = ref_giremi.mismatch.region_mismatch_analysis(chrom, a, b, ds.sam(), ds.fasta(), **kw)
    finally:
        lg.uninstall()
    frames_equal(seam, stock)
