"""Batched form of the reference's per-region analysis (seam ii).

`giremi.mismatch.region_mismatch_analysis` (mismatch.py:345-510) does, for one
footprint: extraction + filters (:367-385), then per strand pair MI -> het
filter -> per-site mean (:387-404), then three DataFrames (:406-510).  Here the
three stages are separate so that the middle one runs ONCE on the GPU for any
number of footprints:

  extract_region(...)        the reference's own get_region_mismatches_with_filters
                             (unchanged, imported from the installed `giremi`) plus
                             the bit-plane encoding of both strands -- worker side
  analyse_extracted(regs)    one submit for every (footprint, strand) unit
                             (api.mi_step_batched, HET_ONLY | SKIP_NONHET), then the frames
  region_mismatch_analysis   drop-in with the reference's signature (one region)
  footprint_bulk_calculation drop-in for giremi/script/giremi.py:20-95: one submit
                             per chunk of footprints instead of two Python MI loops
                             per footprint

The frames have the reference's columns, row order and dtypes; tests compare them
with pandas.testing.assert_frame_equal against the unmodified reference.
Nothing here computes MI on the CPU."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import api
from .encode import concat_plane_batches, encode_mismatches_native

_PAIR_COLUMNS = ['chromosome', 'strand', 'site1_pos', 'site1_type', 'site2_pos', 'site2_type', 'mi']
_SITE_COLUMNS = ['type', 'chromosome', 'strand', 'pos', 'ref', 'change_type', 'ratio', 'allelic_ratio_diff',
                 'depth', 'A:C:T:G', 'up_seq', 'down_seq', 'mean_mi']
_REMOVED_COLUMNS = ['chromosome', 'strand', 'pos', 'removed']
_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}
_STRANDS = ('+', '-')


class RegionExtract:
    """What a worker hands to the parent for one footprint."""

    __slots__ = ("chromosome", "mismatches", "removed", "encoded")

    def __init__(self, chromosome, mismatches, removed):
        self.chromosome = chromosome
        self.mismatches = mismatches            # {'+': {pos: site}, '-': {...}}  (reference objects, untouched)
        self.removed = removed
        # a strand enters the MI step only with at least two sites (mismatch.py:388)
        # (native encoder, csrc/lgmi_host.inl: one PlaneBatch of one unit per strand)
        self.encoded = {s: (encode_mismatches_native(mismatches[s]) if len(mismatches[s]) > 1 else None)
                        for s in _STRANDS}


def extract_region(chromosome, start_pos, end_pos, sam, genome, **filters) -> RegionExtract:
    """Extraction and site filters by the reference's own code (mismatch.py:11-342), then encode."""
    from giremi.mismatch import get_region_mismatches_with_filters
    filters.pop('min_common_reads', None)
    mismatches, removed = get_region_mismatches_with_filters(
        chromosome=chromosome, start_pos=start_pos, end_pos=end_pos, sam=sam, genome=genome, **filters)
    return RegionExtract(chromosome, mismatches, removed)


def _alt_major(site):
    """(ref allele, major non-reference allele, its depth, total depth)  -- mismatch.py:432-440, :452-460."""
    depth = site['depth']
    ref = site['ref']
    total = sum(depth[nt] for nt in depth)
    alts = sorted(([nt, depth[nt]] for nt in depth if nt != ref), key=lambda a: a[1], reverse=True)
    return ref, alts[0][0], alts[0][1], total


def _site_frame_rows(chromosome, strand, sites, mean_of):
    het = [d / t for (_r, _a, d, t) in (_alt_major(s) for s in sites.values() if s['type'] == 'het_snp')]
    allelic_ratio = sum(het) / len(het) if het else 0.5
    rows = []
    for pos, site in sites.items():
        ref, alt, alt_depth, total = _alt_major(site)
        ratio = alt_depth / total
        depth = site['depth']
        acgt = '{}:{}:{}:{}'.format(*(depth[nt] if nt in depth else 0 for nt in 'ACTG'))
        if strand == '+':
            change, up, down = '{}>{}'.format(ref, alt), site['up'], site['down']
        else:
            change = '{}>{}'.format(_COMPLEMENT[ref], _COMPLEMENT[alt])
            up, down = _COMPLEMENT[site['down']], _COMPLEMENT[site['up']]
        if 'N' in change:
            continue                                    # mismatch.py:483-490
        rows.append([site['type'], chromosome, strand, pos, ref, change, ratio, ratio - allelic_ratio, total, acgt,
                     up, down, mean_of.get(pos, np.nan)])
    return rows


class _UnitMeta:
    """positions / types / bad sites of a one-unit PlaneBatch."""

    __slots__ = ("positions", "types", "bad_sites", "n_sites")

    def __init__(self, pb):
        self.positions, self.types, self.bad_sites = pb.positions[0], pb.types[0], pb.bad_sites[0]
        self.n_sites = len(self.positions)


def pair_frame(chromosome, strand, eu, rec):
    """The `.mi.txt` rows of one unit (mismatch.py:407-418, written at giremi.py:400-404) straight
    from the record array: no per-row Python.  Same columns, order and dtypes as the reference's
    DataFrame.from_records over [chromosome, strand, p1, type1, p2, type2, mi] lists."""
    pos = np.asarray(eu.positions, dtype=np.int64)
    typ = np.asarray(eu.types, dtype=object)
    i, j = rec['i'].astype(np.int64), rec['j'].astype(np.int64)
    n = len(rec)
    return pd.DataFrame({
        'chromosome': np.full(n, chromosome, dtype=object), 'strand': np.full(n, strand, dtype=object),
        'site1_pos': pos[i], 'site1_type': typ[i], 'site2_pos': pos[j], 'site2_type': typ[j],
        'mi': rec['mi'].astype(np.float64)}, columns=_PAIR_COLUMNS)


def write_mi_table(path, frames):
    """`.mi.txt` exactly as the CLI writes it (giremi.py:400-404): the pair frames of all regions
    concatenated, tab-separated, no index."""
    frames = [f for f in frames if len(f)]
    df = pd.concat(frames, axis=0) if frames else pd.DataFrame.from_records([], columns=_PAIR_COLUMNS)
    df.to_csv(path, sep='\t', index=False)
    return len(df)


def analyse_extracted(regions, min_common_reads=5, ctx=None):
    """[(df_mismatches, df_mismatch_pair_mi, df_removed_mismatches), ...] for the extracted
    regions, with ONE GPU submit for all of their (footprint, strand) units."""
    units, owner = [], []
    for r, reg in enumerate(regions):
        for s in _STRANDS:
            if reg.encoded[s] is not None:
                units.append(reg.encoded[s])
                owner.append((r, s))
    # only pairs next to a het SNP are evaluated: mismatch.py:393-396 keeps no other pair and
    # mismatch_pair_mi_full is used for nothing else, so the frames are the same
    mode = api.MODE_HET_ONLY | api.MODE_SKIP_NONHET
    res = api.mi_step_batched(concat_plane_batches(units), min_common_reads, mode, ctx=ctx) if units else None
    pair_rows = [{'+': None, '-': None} for _ in regions]
    mean_of = [{'+': {}, '-': {}} for _ in regions]
    site_off = 0
    for u, (r, s) in enumerate(owner):
        eu = _UnitMeta(units[u])
        rec = res.unit_records(u)
        if eu.bad_sites and len(rec):
            bad = np.fromiter(eu.bad_sites, dtype=np.int64)
            if np.isin(rec['i'], bad).any() or np.isin(rec['j'], bad).any():
                raise IndexError('list index out of range')          # mutual_information.py:30/32
        pos = eu.positions
        pair_rows[r][s] = pair_frame(regions[r].chromosome, s, eu, rec)
        mean = res.site_mean[site_off:site_off + eu.n_sites]
        mean_of[r][s] = {p: float(m) for p, m in zip(pos, mean.tolist()) if m == m}   # NaN: in no kept pair
        site_off += eu.n_sites
    out = []
    for r, reg in enumerate(regions):
        parts = [f for f in (pair_rows[r]['+'], pair_rows[r]['-']) if f is not None and len(f)]
        if not parts:
            df_pairs = pd.DataFrame.from_records([], columns=_PAIR_COLUMNS)     # what the reference builds from no rows
        else:
            df_pairs = parts[0] if len(parts) == 1 else pd.concat(parts, axis=0, ignore_index=True)
        site_rows = []
        for s in _STRANDS:
            site_rows += _site_frame_rows(reg.chromosome, s, reg.mismatches[s], mean_of[r][s])
        df_sites = pd.DataFrame.from_records(site_rows, columns=_SITE_COLUMNS)
        df_removed = pd.DataFrame.from_records(
            [[reg.chromosome, s, pos, reg.removed[s][pos]['removed']] for s in _STRANDS for pos in reg.removed[s]],
            columns=_REMOVED_COLUMNS)
        out.append((df_sites, df_pairs, df_removed))
    return out


def region_mismatch_analysis(chromosome, start_pos, end_pos, sam, genome, min_common_reads=5, **filters):
    """Drop-in for giremi.mismatch.region_mismatch_analysis (mismatch.py:345-510): same
    arguments, same three DataFrames."""
    reg = extract_region(chromosome, start_pos, end_pos, sam, genome, **filters)
    return analyse_extracted([reg], min_common_reads)[0]


_FILTER_KEYS = ('keep_non_spliced_read', 'min_dist_from_splice', 'min_allele_depth', 'min_allele_ratio',
                'min_total_depth', 'homopoly_length', 'min_het_snp_ratio', 'max_het_snp_ratio',
                'mismatch_window_size', 'max_window_mismatch', 'max_window_mismatch_type', 'mode')


def footprint_bulk_calculation(footprints, variables):
    """Drop-in for giremi.script.giremi.footprint_bulk_calculation (giremi.py:20-95):
    same inputs, same four DataFrames; every footprint of the chunk is extracted first
    and the MI step of all of them is one GPU submit."""
    import pysam
    from giremi.fileio import read_simple_repeat_intervals, read_snp_positions_in_region
    from giremi.strand import correct_read_strand_in_region
    sam = pysam.AlignmentFile(variables['bam_file'], 'rb')
    genome = pysam.FastaFile(variables['genome_file'])
    vcf = pysam.VariantFile(variables['snp_file'])
    gtf = pysam.TabixFile(variables['gtf_file'])
    repeats = read_simple_repeat_intervals(variables['repeat_file'])
    strand_list, regions = [], []
    for chromosome, start_pos, end_pos, _rc in footprints:
        read_strand_dict = None
        if not variables['skip_strand_correction']:
            read_strand_list = correct_read_strand_in_region(
                chromosome, start_pos, end_pos, sam, gtf, genome, variables['gene_padding'], variables['exon_padding'],
                keep_non_spliced_read=variables['keep_non_spliced_read'], mode=variables['mode'])
            strand_list.extend(read_strand_list)
            read_strand_dict = dict([rname, corrected] for rname, _old, corrected in read_strand_list)
        snp_positions = read_snp_positions_in_region(vcf, chromosome, start_pos, end_pos)
        simple_repeat_intervals = [[rs, re] for rs, re in repeats[chromosome]
                                   if (rs > end_pos) or (re < start_pos)]          # as written at giremi.py:55-59
        regions.append(extract_region(
            chromosome, start_pos, end_pos, sam, genome, simple_repeat_intervals=simple_repeat_intervals,
            snp_positions=snp_positions, read_strand_dict=read_strand_dict,
            **{k: variables[k] for k in _FILTER_KEYS}))
    frames = analyse_extracted(regions, variables['mi_min_common_reads'])
    mismatch_df = pd.concat([f[0] for f in frames], axis=0)
    mi_df = pd.concat([f[1] for f in frames], axis=0)
    removed_df = pd.concat([f[2] for f in frames], axis=0)
    strand_df = pd.DataFrame.from_records(strand_list,
                                          columns=['read_name', 'original_read_strand', 'corrected_read_strand'])
    for fh in (sam, genome, vcf, gtf):
        fh.close()
    return mismatch_df, mi_df, strand_df, removed_df
