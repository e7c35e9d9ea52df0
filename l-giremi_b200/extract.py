"""Per-region extraction straight to bit-planes (f1 + f2; north_star part 1).

`get_region_sites(...)` produces what giremi.mismatch.get_region_mismatches_with_filters
(mismatch.py:11-342) produces -- the same surviving sites with the same fields, the same removed
sites with the same reasons, in the same order -- for `mode='cs'`, with two differences in HOW:

  * every read's cs tag goes through the C++ scanner lgmi_cs_scan (csrc/lgmi_host.inl: tokenizer,
    contig coordinates, introns and the splice-distance FILTER 1 of mismatch.py:99-141 in one
    pass) instead of CS.from_cs_tag_string + get_mismatches + get_introns + two interval helpers;
  * read names are interned once per region: a site's allele lists hold small integers, and the
    unit's bit-planes are written from them directly (`encode_indexed`: one fancy-index store per
    allele list) -- no read-name strings, no per-pair dict rebuild, no flattening of a dict into
    name blobs.

The site filters themselves are the reference's, step for step (each helper cites its lines),
including what the reference does by accident and the output tables show: an (empty) entry for
every pileup column (mismatch.py:166), and sites dropped by the window filter coming back empty
when a later site's window touches them (:228, a defaultdict read) to be dropped again as
'too few usable reads after filters'.  tests/test_extract_host.py compares both outputs against
the unmodified reference on simulated reads, field by field."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .encode import PlaneBatch, row_words
from ._lib import SITE_HAS_OTHER, SITE_TYPE_CODE, UNIT_DESC

_STRANDS = ('+', '-')
_PAIR = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def _new_site(removed=False):
    site = {'ref': '', 'type': 'mismatch', 'depth': {}, 'nt': {}, 'neighbor': {}, 'up': '', 'down': ''}
    if removed:
        site['removed'] = ''
    return site


class _Sites(dict):
    """pos -> site with the reference's defaultdict behaviour: READING a missing position creates it."""

    def __missing__(self, pos):
        site = self[pos] = _new_site()
        return site


class CsScanner:
    """lgmi_cs_scan with reusable output buffers (one call per read)."""

    def __init__(self, cap=256):
        self._lib = _lib.load()
        self._alloc(cap, cap)

    def _alloc(self, cap, cap_i):
        self.cap, self.cap_i = cap, cap_i
        self.pos = np.empty(cap, np.int64)
        self.ref = np.empty(cap, "S1")
        self.alt = np.empty(cap, "S1")
        self.lo = np.empty(cap_i, np.int64)
        self.hi = np.empty(cap_i, np.int64)
        self.n, self.ni = C.c_uint32(), C.c_uint32()

    def scan(self, cs_tag, reference_start, min_dist_from_splice):
        """([(pos, ref base, read base)] after the splice-distance filter, number of introns)."""
        raw = cs_tag.encode() if isinstance(cs_tag, str) else bytes(cs_tag)
        while True:
            rc = self._lib.lgmi_cs_scan(raw, len(raw), int(reference_start), int(min_dist_from_splice), self.cap,
                                        _lib.ptr(self.pos), _lib.ptr(self.ref), _lib.ptr(self.alt), C.byref(self.n),
                                        self.cap_i, _lib.ptr(self.lo), _lib.ptr(self.hi), C.byref(self.ni))
            if rc == -3:                                    # LGMI_ERR_NOMEM: counts are set, grow and rescan
                self._alloc(max(self.cap, self.n.value + 1), max(self.cap_i, self.ni.value + 1))
                continue
            if rc != 0:
                raise KeyError("cs tag not understood (liblgmi error %d): %r" % (rc, cs_tag[:60]))
            break
        n = self.n.value
        return list(zip(self.pos[:n].tolist(), self.ref[:n].tolist(), self.alt[:n].tolist())), self.ni.value


def _collect_reads(sam, chromosome, start_pos, end_pos, keep_non_spliced_read, min_dist_from_splice, read_strand_dict,
                   sites, index_of, names):
    """mismatch.py:66-147: every read's substitutions (after FILTER 1) into sites[strand][pos]."""
    scanner = CsScanner()
    for read in sam.fetch(chromosome, start_pos, end_pos):
        name = read.query_name
        strand = '-' if read.is_reverse else '+'
        if name in read_strand_dict:
            strand = read_strand_dict[name]
        else:
            read_strand_dict[name] = strand
        found, n_introns = scanner.scan(read.get_tag('cs'), read.reference_start, min_dist_from_splice)
        if not keep_non_spliced_read and n_introns == 0:
            continue
        if not found:
            continue
        k = index_of.get(name)
        if k is None:
            k = index_of[name] = len(names)
            names.append(name)
        table = sites[strand]
        for pos, ref, alt in found:                         # ascending position: the scanner emits them in tag order
            site = table[pos]
            site['ref'] = ref.decode()
            site['nt'].setdefault(alt.decode(), []).append(k)


def _add_reference_reads(sam, chromosome, start_pos, end_pos, strand, table, read_strand_dict, index_of, names):
    """mismatch.py:157-188: reads carrying the reference base at every site seen so far; every pileup column
    is looked up in the table (and so enters it, empty, as in the reference)."""
    wanted = set(table)
    for column in sam.pileup(contig=chromosome, start=start_pos, stop=end_pos):
        pos = column.pos
        ref = table[pos]['ref']
        if pos not in wanted or ref.upper() not in _PAIR:
            continue
        got = []
        for name, base in zip(column.get_query_names(), column.get_query_sequences()):
            if base.upper() == ref and read_strand_dict[name] == strand:
                k = index_of.get(name)
                if k is None:
                    k = index_of[name] = len(names)
                    names.append(name)
                got.append(k)
        table[pos]['nt'].setdefault(ref, []).extend(got)


def _set_depth(table):
    """mismatch.py:203-208 / :316-323."""
    for pos in sorted(table):
        site = table[pos]
        for nt, reads in site['nt'].items():
            site['depth'][nt] = len(reads)


def _reset_depth(table):
    for pos in sorted(table):
        site = table[pos]
        site['depth'] = {nt: len(reads) for nt, reads in site['nt'].items()}


def _drop(table, removed, pos, reason):
    removed[pos] = table.pop(pos)
    removed[pos]['removed'] = reason


def _window_filter(table, removed, strand, window, max_n, max_types):
    """mismatch.py:210-240.  `positions` is fixed before the loop, so a site dropped earlier in the loop is
    READ again through the defaultdict when a later window covers it and comes back as an empty entry."""
    half = round(window / 2)
    positions = sorted(table)
    for pos in positions:
        near = [a for a in positions if a != pos and pos - half <= a < pos + half]
        if not near:
            continue
        site = table[pos]
        for a in near:
            other = table[a]
            ref = other['ref']
            for nt in other['depth']:
                if nt == ref:
                    continue
                change = '{}>{}'.format(ref, nt) if strand == '+' else '{}>{}'.format(_PAIR[ref], _PAIR[nt])
                site['neighbor'][change] = site['neighbor'].get(change, 0) + 1
        if sum(site['neighbor'].values()) > max_n and len(site['neighbor']) > max_types:
            _drop(table, removed, pos, 'too many window mismatches')


def _allele_filters(table, removed, min_allele_depth, min_allele_ratio, min_total_depth):
    """mismatch.py:242-282 (depth keeps the dropped alleles' counts until the final recount, as there)."""
    for pos in sorted(table):
        site = table[pos]
        for nt in list(site['nt']):
            if site['depth'][nt] < min_allele_depth:
                site['nt'].pop(nt)
    for pos in sorted(table):
        site = table[pos]
        total = sum(site['depth'].values())
        for nt in list(site['nt']):
            if site['depth'][nt] / total < min_allele_ratio:
                site['nt'].pop(nt)
    for pos in sorted(table):
        if sum(table[pos]['depth'].values()) < min_total_depth:
            _drop(table, removed, pos, 'too few usable reads after filters')
    for pos in sorted(table):
        if len(table[pos]['nt']) < 2:
            _drop(table, removed, pos, 'not enough allele after filters')


def _sequence_filters(table, removed, genome, chromosome, homopoly_length, simple_repeat_intervals):
    """mismatch.py:283-313: flanking bases, homopolymers, simple repeats (membership start <= pos < end over the
    intervals sorted by start, utils.py:20-31)."""
    half = int(homopoly_length / 2)
    for pos in sorted(table):
        site = table[pos]
        left = genome.fetch(chromosome, pos - homopoly_length, pos).upper()
        right = genome.fetch(chromosome, pos + 1, pos + homopoly_length + 1).upper()
        site['up'], site['down'] = left[-1].upper(), right[0].upper()
        if len(set(left)) == 1 or len(set(right)) == 1 or len(set(left[-half:] + right[0:half])) == 1:
            _drop(table, removed, pos, 'in homopoly regions')
    positions = sorted(table)
    if positions:
        simple_repeat_intervals.sort(key=lambda a: a[0])                  # in place, as the reference's helper does
        starts = [a for a, _b in simple_repeat_intervals]
        ends = [b for _a, b in simple_repeat_intervals]
        inside = np.searchsorted(starts, positions, side='right') - np.searchsorted(ends, positions, side='right') == 1
        for pos, hit in zip(positions, np.atleast_1d(inside).tolist()):
            if hit:
                _drop(table, removed, pos, 'in simple repeat regions')


def _mark_snps(table, snp_positions, lo, hi):
    """mismatch.py:324-340."""
    for pos in sorted(table):
        if pos in snp_positions:
            site = table[pos]
            total = sum(site['depth'].values())
            major = max(d / total for d in site['depth'].values())
            site['type'] = 'het_snp' if lo <= major <= hi else 'snp'


def get_region_sites(chromosome, start_pos, end_pos, sam, genome, keep_non_spliced_read=False, min_dist_from_splice=4,
                     min_allele_depth=3, min_allele_ratio=0.1, min_total_depth=6, homopoly_length=5,
                     simple_repeat_intervals=[], snp_positions=[], read_strand_dict=None, min_het_snp_ratio=0.35,
                     max_het_snp_ratio=0.65, mismatch_window_size=100, max_window_mismatch=10,
                     max_window_mismatch_type=3, mode='cs'):
    """(sites, removed, read names): the two dicts of get_region_mismatches_with_filters (same signature) with
    read INDICES in the 'nt' lists; names[k] is the name of read k."""
    if mode != 'cs':
        raise ValueError("get_region_sites reads cs tags; use the reference's extraction for mode=%r" % mode)
    sites = {s: _Sites() for s in _STRANDS}
    removed = {s: {} for s in _STRANDS}
    if read_strand_dict is None:
        read_strand_dict = {}
    index_of, names = {}, []
    _collect_reads(sam, chromosome, start_pos, end_pos, keep_non_spliced_read, min_dist_from_splice, read_strand_dict,
                   sites, index_of, names)
    for strand in _STRANDS:
        table = sites[strand]
        if not table:
            continue
        _add_reference_reads(sam, chromosome, start_pos, end_pos, strand, table, read_strand_dict, index_of, names)
        _set_depth(table)
        _window_filter(table, removed[strand], strand, mismatch_window_size, max_window_mismatch, max_window_mismatch_type)
        _allele_filters(table, removed[strand], min_allele_depth, min_allele_ratio, min_total_depth)
        _sequence_filters(table, removed[strand], genome, chromosome, homopoly_length, simple_repeat_intervals)
        _reset_depth(table)
        _mark_snps(table, snp_positions, min_het_snp_ratio, max_het_snp_ratio)
    return sites, removed, names


def encode_indexed(table) -> PlaneBatch:
    """One unit's bit-planes from sites whose 'nt' lists hold read indices.  The dict semantics the kernels
    cannot see are applied here: a read listed twice at a site keeps its LAST allele in ('nt' order, list order)
    (mutual_information.py:15-16: later stores overwrite earlier ones), major / minor by `depth` descending with
    the stable tie-break (:25-32), every other allele -> "other" (:33-38).  Reads are renumbered densely over the
    unit (the MI of a pair does not depend on the order of the reads)."""
    positions = sorted(table)
    S = len(positions)
    used = np.unique(np.fromiter((k for pos in positions for reads in table[pos]['nt'].values() for k in reads),
                                 dtype=np.int64)) if S else np.zeros(0, np.int64)
    R = len(used)
    W = row_words(R)
    bits = np.zeros((S, 3, W * 32), dtype=np.uint8)
    flags = np.zeros(S, dtype=np.uint8)
    types, bad = [], []
    for s, pos in enumerate(positions):
        site = table[pos]
        types.append(site['type'])
        ranked = sorted(site['depth'].items(), key=lambda kv: -kv[1])       # stable
        major = ranked[0][0] if ranked else None
        minor = ranked[1][0] if len(ranked) > 1 else None
        if len(ranked) < 2:
            bad.append(s)
        label = np.full(R, 255, dtype=np.uint8)
        for allele, reads in site['nt'].items():
            if reads:
                code = 2 if allele == major else (1 if allele == minor else 0)
                label[np.searchsorted(used, np.asarray(reads, dtype=np.int64))] = code
        bits[s, 0, :R] = label == 2
        bits[s, 1, :R] = label == 1
        bits[s, 2, :R] = label != 255
        flags[s] = SITE_TYPE_CODE[site['type']] | (SITE_HAS_OTHER if (label == 0).any() else 0)
    planes = np.ascontiguousarray(np.packbits(bits, axis=-1, bitorder='little')).view('<u4').reshape(-1)
    units = np.zeros(1, dtype=UNIT_DESC)
    units[0] = (0, S, R, W, 0)
    return PlaneBatch(units, planes, flags, [positions], [types], [frozenset(bad)])
