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
  main                       drop-in for giremi/script/giremi.py:324-455 (the `l-giremi`
                             console script): the pool workers only extract and encode
                             (extract_footprints); the PARENT owns the GPUs -- every
                             visible one, units partitioned by pair-count cost
                             (multigpu.DevicePool) -- and the mip / label pass
                             (:415-429) is one launch instead of a Python call per row

The frames have the reference's columns, row order and dtypes; tests compare them
with pandas.testing.assert_frame_equal against the unmodified reference.
Nothing here computes MI on the CPU."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import api
from .encode import concat_plane_batches, encode_mismatches_native
from .extract import encode_indexed, get_region_sites

_PAIR_COLUMNS = ['chromosome', 'strand', 'site1_pos', 'site1_type', 'site2_pos', 'site2_type', 'mi']
_SITE_COLUMNS = ['type', 'chromosome', 'strand', 'pos', 'ref', 'change_type', 'ratio', 'allelic_ratio_diff',
                 'depth', 'A:C:T:G', 'up_seq', 'down_seq', 'mean_mi']
_REMOVED_COLUMNS = ['chromosome', 'strand', 'pos', 'removed']
_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}
_STRANDS = ('+', '-')


class RegionExtract:
    """What a worker hands to the parent for one footprint: the bit-planes of its (up to two)
    units and everything of the three output tables that does not depend on the MI step.  The
    reference's dicts (read-name lists) stay in the worker unless keep_dicts is set."""

    __slots__ = ("chromosome", "encoded", "site_rows", "site_pos", "removed_rows", "mismatches", "removed")

    def __init__(self, chromosome, mismatches, removed, keep_dicts=True, indexed=False):
        self.chromosome = chromosome
        # a strand enters the MI step only with at least two sites (mismatch.py:388)
        # indexed: the allele lists hold read indices (extract.get_region_sites) and the planes are written from
        # them directly; else read names through the native dict encoder (csrc/lgmi_host.inl, lgmi_encode_unit)
        encode = encode_indexed if indexed else encode_mismatches_native
        self.encoded = {s: (encode(mismatches[s]) if len(mismatches[s]) > 1 else None) for s in _STRANDS}
        self.site_rows, self.site_pos = {}, {}
        for s in _STRANDS:
            self.site_rows[s], self.site_pos[s] = _site_frame_rows(chromosome, s, mismatches[s])
        self.removed_rows = [[chromosome, s, pos, removed[s][pos]['removed']] for s in _STRANDS for pos in removed[s]]
        self.mismatches = mismatches if keep_dicts else None   # {'+': {pos: site}, '-': {...}}  (reference objects)
        self.removed = removed if keep_dicts else None


def native_extraction_enabled():
    """LGMI_NATIVE_EXTRACT=0 keeps the reference's own extraction inside the batched seam."""
    import os
    return os.environ.get("LGMI_NATIVE_EXTRACT", "1") != "0"


def extract_region(chromosome, start_pos, end_pos, sam, genome, keep_dicts=True, native=None, **filters) -> RegionExtract:
    """Extraction, site filters and encoding of one footprint.  For cs-tag input (the reference's default mode)
    the reads go through the C++ cs scanner and the planes are written from interned read indices
    (extract.get_region_sites, the same sites as mismatch.py:11-342 field for field); for MD / CIGAR input
    (mode='cigar') or native=False, the reference's own function runs and its dicts are encoded."""
    filters.pop('min_common_reads', None)
    if native is None:
        native = native_extraction_enabled()
    if native and filters.get('mode', 'cs') == 'cs':
        sites, removed, _names = get_region_sites(chromosome, start_pos, end_pos, sam, genome, **filters)
        return RegionExtract(chromosome, sites, removed, keep_dicts=keep_dicts, indexed=True)
    from giremi.mismatch import get_region_mismatches_with_filters
    mismatches, removed = get_region_mismatches_with_filters(
        chromosome=chromosome, start_pos=start_pos, end_pos=end_pos, sam=sam, genome=genome, **filters)
    return RegionExtract(chromosome, mismatches, removed, keep_dicts=keep_dicts)


def _alt_major(site):
    """(ref allele, major non-reference allele, its depth, total depth)  -- mismatch.py:432-440, :452-460."""
    depth = site['depth']
    ref = site['ref']
    total = sum(depth[nt] for nt in depth)
    alts = sorted(([nt, depth[nt]] for nt in depth if nt != ref), key=lambda a: a[1], reverse=True)
    return ref, alts[0][0], alts[0][1], total


def _site_frame_rows(chromosome, strand, sites):
    """Rows of the site table (mismatch.py:420-494) with mean_mi still NaN, and the position of each
    row (the parent fills mean_mi from the MI step by position)."""
    het = [d / t for (_r, _a, d, t) in (_alt_major(s) for s in sites.values() if s['type'] == 'het_snp')]
    allelic_ratio = sum(het) / len(het) if het else 0.5
    rows, positions = [], []
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
                     up, down, np.nan])
        positions.append(pos)
    return rows, positions


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


def _check_bad_sites(units, metas, min_common_reads, ctx):
    """The reference raises IndexError (mutual_information.py:30/32) as soon as a pair with enough common
    reads touches a site whose `depth` holds fewer than two alleles -- any such pair, not only those next to a
    het SNP.  The batched step evaluates het-adjacent pairs only, so units with such a site (the reference's
    filters never let one through, mismatch.py:275-282) are looked at once more with every pair kept."""
    flagged = [k for k, m in enumerate(metas) if m.bad_sites]
    if not flagged:
        return
    res = api.mi_step_batched(concat_plane_batches([units[k] for k in flagged]), min_common_reads,
                              api.MODE_ALL_PAIRS, ctx=ctx)
    for n, k in enumerate(flagged):
        rec = res.unit_records(n)
        bad = np.fromiter(metas[k].bad_sites, dtype=np.int64)
        if len(rec) and (np.isin(rec['i'], bad).any() or np.isin(rec['j'], bad).any()):
            raise IndexError('list index out of range')


def analyse_extracted(regions, min_common_reads=5, ctx=None, pool=None):
    """[(df_mismatches, df_mismatch_pair_mi, df_removed_mismatches), ...] for the extracted
    regions, with ONE GPU submit for all of their (footprint, strand) units -- or, given a
    multigpu.DevicePool, one submit per GPU over a cost-balanced partition of the units."""
    units, owner = [], []
    for r, reg in enumerate(regions):
        for s in _STRANDS:
            if reg.encoded[s] is not None:
                units.append(reg.encoded[s])
                owner.append((r, s))
    # only pairs next to a het SNP are evaluated: mismatch.py:393-396 keeps no other pair and
    # mismatch_pair_mi_full is used for nothing else, so the frames are the same
    mode = api.MODE_HET_ONLY | api.MODE_SKIP_NONHET
    res = None
    metas = [_UnitMeta(u) for u in units]
    if units:
        first = pool.contexts[0] if pool is not None else ctx
        _check_bad_sites(units, metas, min_common_reads, first)
        pb = concat_plane_batches(units)
        import time
        t0 = time.perf_counter()
        res = pool.run(pb, min_common_reads, mode) if pool is not None else \
            api.mi_step_batched(pb, min_common_reads, mode, ctx=ctx)
        last_step_times['submit_s'] = time.perf_counter() - t0
    pair_rows = [{'+': None, '-': None} for _ in regions]
    mean_of = [{'+': {}, '-': {}} for _ in regions]
    site_off = 0
    for u, (r, s) in enumerate(owner):
        eu = metas[u]
        rec = res.unit_records(u)
        pair_rows[r][s] = pair_frame(regions[r].chromosome, s, eu, rec)
        mean = res.site_mean[site_off:site_off + eu.n_sites]
        mean_of[r][s] = {p: float(m) for p, m in zip(eu.positions, mean.tolist()) if m == m}   # NaN: in no kept pair
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
            found = mean_of[r][s]
            for row, pos in zip(reg.site_rows[s], reg.site_pos[s]):
                site_rows.append(row[:-1] + [found.get(pos, np.nan)])
        df_sites = pd.DataFrame.from_records(site_rows, columns=_SITE_COLUMNS)
        df_removed = pd.DataFrame.from_records(reg.removed_rows, columns=_REMOVED_COLUMNS)
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


def extract_footprints(footprints, variables, keep_dicts=False):
    """The worker half of giremi/script/giremi.py:20-95 for one chunk of footprints: strand
    correction, SNP positions, repeats, extraction + filters (all the reference's own code) and the
    bit-plane encoding.  No MI, no CUDA: safe in a forked pool worker.  Returns
    (read-strand rows, [RegionExtract, ...])."""
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
            chromosome, start_pos, end_pos, sam, genome, keep_dicts=keep_dicts,
            simple_repeat_intervals=simple_repeat_intervals,
            snp_positions=snp_positions, read_strand_dict=read_strand_dict,
            **{k: variables[k] for k in _FILTER_KEYS}))
    for fh in (sam, genome, vcf, gtf):
        fh.close()
    return strand_list, regions


def _chunk_frames(strand_list, frames):
    """The four DataFrames footprint_bulk_calculation returns for one chunk (giremi.py:79-93)."""
    mismatch_df = pd.concat([f[0] for f in frames], axis=0)
    mi_df = pd.concat([f[1] for f in frames], axis=0)
    removed_df = pd.concat([f[2] for f in frames], axis=0)
    strand_df = pd.DataFrame.from_records(strand_list,
                                          columns=['read_name', 'original_read_strand', 'corrected_read_strand'])
    return mismatch_df, mi_df, strand_df, removed_df


def footprint_bulk_calculation(footprints, variables):
    """Drop-in for giremi.script.giremi.footprint_bulk_calculation (giremi.py:20-95):
    same inputs, same four DataFrames; every footprint of the chunk is extracted first
    and the MI step of all of them is one GPU submit.  Called inside a pool worker (the stock
    `main`), the worker's GPU is worker index modulo the number of GPUs (api.default_device)."""
    strand_list, regions = extract_footprints(footprints, variables)
    return _chunk_frames(strand_list, analyse_extracted(regions, variables['mi_min_common_reads']))


# argparse destination -> key of the `variables` dict the reference's main builds (giremi.py:327-352)
_VARIABLE_OF_ARG = {
    'bam_file': 'bam_file', 'genome_fasta': 'genome_file', 'annotation_gtf': 'gtf_file', 'repeat_txt': 'repeat_file',
    'snp_bcf': 'snp_file', 'padding_exon': 'exon_padding', 'padding_gene': 'gene_padding',
    'homopoly_length': 'homopoly_length', 'keep_non_spliced_read': 'keep_non_spliced_read',
    'max_het_snp_ratio': 'max_het_snp_ratio', 'min_allele_depth': 'min_allele_depth',
    'min_allele_ratio': 'min_allele_ratio', 'min_dist_from_splice': 'min_dist_from_splice',
    'min_het_snp_ratio': 'min_het_snp_ratio', 'min_total_depth': 'min_total_depth',
    'mi_p_threshold': 'mip_threshold', 'mi_min_common_read': 'mi_min_common_reads',
    'mi_calculation_only': 'mi_calculation_only', 'mismatch_window_size': 'mismatch_window_size',
    'max_window_mismatch': 'max_window_mismatch', 'max_window_mismatch_type': 'max_window_mismatch_type',
    'mode': 'mode', 'model': 'model', 'skip_strand_correction': 'skip_strand_correction',
}

# wall-clock seconds of the last main(): what tools/run_cli.py reports as the MI-step time of the patched CLI
last_run_times = {}
# ... and of the last analyse_extracted(): the GPU submit alone (encode and frames excluded)
last_step_times = {}


def main():
    """Drop-in for giremi.script.giremi.main (giremi.py:324-455), the `l-giremi` console script:
    same arguments (the reference's own parse_args), same five output files.

    Differences in HOW, none in WHAT: (1) the `-t` pool workers extract and encode only
    (extract_footprints) -- no worker touches CUDA; (2) the parent, after the pool has returned,
    runs the MI step of ALL chunks' units over every visible GPU (multigpu.DevicePool: LPT
    partition by pair-count cost, one submit per GPU, results merged back in the reference's
    order); (3) the empirical p-value of every site (giremi.py:415-429: ecdf over the het-SNP
    means, then a Python call per row) is one lgmi_ecdf launch (api.mip_and_calls).  The GLM
    scoring (run_glm) is the reference's own."""
    import logging
    import multiprocessing as mp
    import time
    from functools import partial
    import giremi.script.giremi as cli
    from giremi.footprint import get_footprints
    from . import multigpu
    args = cli.parse_args()
    variables = {key: getattr(args, dest) for dest, key in _VARIABLE_OF_ARG.items()}
    logging.basicConfig(format='%(asctime)s %(levelname)s %(message)s', level=logging.INFO)
    logging.info('Get regions that are covered by enough reads.')
    footprints = get_footprints(variables['bam_file'], args.chromosomes, variables['min_total_depth'])
    n = int(len(footprints) / args.thread / 2)                         # chunking as giremi.py:367-370
    chunks = [footprints[i:(i + n)] for i in range(0, len(footprints), n)]
    logging.info('Calculate mismatches in each region.')
    t0 = time.perf_counter()
    with mp.Pool(args.thread) as p:                                     # (the workers are forked here ...)
        pending = p.map_async(partial(extract_footprints, variables=variables), chunks)
        # ... so CUDA can start in the parent while they extract: context creation, module load, ln table
        pool = multigpu.get_pool()
        t_cuda = time.perf_counter() - t0
        extracted = pending.get()
    t1 = time.perf_counter()
    regions = [reg for _strand_rows, regs in extracted for reg in regs]
    frames = analyse_extracted(regions, variables['mi_min_common_reads'], pool=pool)
    t2 = time.perf_counter()
    results, k = [], 0
    for strand_rows, regs in extracted:
        results.append(_chunk_frames(strand_rows, frames[k:k + len(regs)]))
        k += len(regs)
    mismatch_df = pd.concat([r[0] for r in results], axis=0)
    mi_df = pd.concat([r[1] for r in results], axis=0)
    strand_df = pd.concat([r[2] for r in results], axis=0)
    removed_df = pd.concat([r[3] for r in results], axis=0)
    strand_df.to_csv(args.output_prefix + '.strand.txt', sep='\t', index=False)
    mi_df.to_csv(args.output_prefix + '.mi.txt', sep='\t', index=False)
    removed_df.to_csv(args.output_prefix + '.removed.txt', sep='\t', index=False)
    t3 = time.perf_counter()
    mip_s = 0.0
    if not variables['mi_calculation_only']:
        logging.info('Score the RNA editing sites.')
        mismatch_df.loc[:, 'mip'] = np.nan
        if mismatch_df['mean_mi'].notna().sum() > 0:
            ta = time.perf_counter()
            mip, _call = api.mip_and_calls(mismatch_df['mean_mi'].to_numpy(dtype=np.float64),
                                           mismatch_df['type'].to_numpy(), variables['mip_threshold'],
                                           ctx=pool.contexts[0])
            mismatch_df.loc[:, 'mip'] = mip
            mip_s = time.perf_counter() - ta
        glmresult, scorepf = cli.run_glm(mismatch_df, mip_threshold=variables['mip_threshold'], model=variables['model'])
        del glmresult['label']
        del glmresult['id']
        glmresult.to_csv(args.output_prefix + '.mismatch.txt', sep='\t', index=False)
        scorepf.to_csv(args.output_prefix + '.score_performance.txt', sep='\t', index=False)
    last_run_times.update({'extract_pool_s': t1 - t0, 'cuda_start_s': t_cuda, 'mi_step_s': t2 - t1,
                           'gpu_submit_s': last_step_times.get('submit_s'), 'write_tables_s': t3 - t2, 'mip_s': mip_s,
                           'n_gpus': len(pool), 'n_units': sum(1 for reg in regions for s in _STRANDS
                                                               if reg.encoded[s] is not None)})
    logging.info('All done!')
