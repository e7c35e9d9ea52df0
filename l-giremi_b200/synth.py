"""Synthetic read x site allele data for the MI step.

One generator, two encodings: every unit can be rendered as the reference's
``mismatches`` dict (for the CPU baseline / oracle) and as bit-planes (for the
GPU path), from the same draw.

Model per unit (footprint x strand) with S candidate sites and R long reads:
  * haplotype h_r ~ Bernoulli(0.5) per read;
  * a read covers one contiguous run of sites; run length uniform so that the
    mean covered fraction is `cov`, start uniform among valid starts;
  * site kinds: 10 % het_snp (allele follows the haplotype, 1 % flips),
    5 % snp (alt frequency 0.9), 85 % mismatch with edit level ~ Beta(2,5),
    30 % of those linked to the haplotype (alt only on h_r = 1);
  * a third allele with probability 0.005 per covered read (label "other");
  * every site is forced to have >= 2 covered reads and >= 2 alleles, as the
    reference's filters guarantee (mismatch.py:275-282);
  * 'nt' insertion order = alt alleles by first occurrence in read order, the
    reference allele last (mismatch.py:144-147, :188); depth = list length.
Positions are 1000 + 37*s; read names "r%07d".
"""
from __future__ import annotations

import numpy as np

from ._lib import SITE_HET_SNP, SITE_MISMATCH, SITE_SNP, SITE_TYPE_NAMES, UNIT_DESC
from .encode import UNCOVERED, EncodedUnit, PlaneBatch, pack_labels, row_words, site_flag_bytes

REF, ALT, THIRD, NOCOV = 0, 1, 2, 255     # allele codes in the raw draw
ALLELE_LETTER = {REF: 'A', ALT: 'G', THIRD: 'T'}


def _run_length_bounds(S, cov):
    """Uniform run-length range [lo, hi] with mean ~= cov*S."""
    mean = min(max(cov * S, 1.0), float(S))
    if mean <= (S + 1) / 2:
        lo, hi = 1, max(1, int(round(2 * mean - 1)))
    else:
        hi, lo = S, min(S, int(round(2 * mean - S)))
    return lo, min(hi, S)


def draw_alleles(rng, G, S, R, cov):
    """Raw draw for G units of identical shape: (alleles uint8 [G,S,R], kinds uint8 [G,S])."""
    lo, hi = _run_length_bounds(S, cov)
    length = rng.integers(lo, hi + 1, size=(G, R))
    start = (rng.random((G, R)) * (S - length + 1)).astype(np.int64)
    s_idx = np.arange(S).reshape(1, S, 1)
    covered = (s_idx >= start[:, None, :]) & (s_idx < (start + length)[:, None, :])
    hap = rng.random((G, 1, R)) < 0.5
    kind_u = rng.random((G, S))
    kinds = np.where(kind_u < 0.10, SITE_HET_SNP, np.where(kind_u < 0.15, SITE_SNP, SITE_MISMATCH)).astype(np.uint8)
    edit = rng.beta(2.0, 5.0, size=(G, S, 1))
    linked = rng.random((G, S, 1)) < 0.30
    u = rng.random((G, S, R), dtype=np.float32)
    k3 = kinds[:, :, None]
    alt_het = hap ^ (u < 0.01)
    alt_snp = u < 0.9
    alt_mm = (u < edit) & (~linked | hap)
    alt = np.where(k3 == SITE_HET_SNP, alt_het, np.where(k3 == SITE_SNP, alt_snp, alt_mm))
    alleles = np.where(alt, ALT, REF).astype(np.uint8)
    third = rng.random((G, S, R), dtype=np.float32) < 0.005
    alleles[third] = THIRD
    alleles[~covered] = NOCOV
    _repair(rng, alleles)
    return alleles, kinds


def _repair(rng, alleles):
    """Force >= 2 covered reads and >= 2 distinct alleles at every site."""
    G, S, R = alleles.shape
    ncov = (alleles != NOCOV).sum(axis=2)
    for g, s in zip(*np.nonzero(ncov < 2)):
        need = 2 - int(ncov[g, s])
        free = np.nonzero(alleles[g, s] == NOCOV)[0]
        pick = rng.choice(free, size=min(need, len(free)), replace=False)
        alleles[g, s, pick] = REF
    n_ref = (alleles == REF).sum(axis=2)
    n_alt = (alleles == ALT).sum(axis=2)
    n_thr = (alleles == THIRD).sum(axis=2)
    mono = ((n_ref > 0).astype(int) + (n_alt > 0) + (n_thr > 0)) < 2
    for g, s in zip(*np.nonzero(mono)):
        cov_idx = np.nonzero(alleles[g, s] != NOCOV)[0]
        r = int(rng.choice(cov_idx))
        alleles[g, s, r] = ALT if alleles[g, s, r] != ALT else REF


def labels_from_alleles(alleles):
    """Raw alleles -> label matrix (2 major / 1 minor / 0 other / UNCOVERED),
    ranking by site-wide depth, ties by 'nt' insertion order (alts by first
    occurrence in read order, reference allele last)."""
    shape = alleles.shape
    a = alleles.reshape(-1, shape[-1])
    n = np.stack([(a == REF).sum(1), (a == ALT).sum(1), (a == THIRD).sum(1)], axis=1).astype(np.int64)
    R = shape[-1]
    first_alt = np.where((a == ALT).any(1), (a == ALT).argmax(1), R + 1)
    first_thr = np.where((a == THIRD).any(1), (a == THIRD).argmax(1), R + 1)
    # insertion rank: smaller = earlier in the dict; the reference allele is last
    ins = np.stack([np.full(len(a), 2 * R + 5), first_alt, first_thr], axis=1)
    # sort key: depth descending, then insertion rank ascending; absent alleles never rank
    key = -n * (4 * R + 16) + ins
    key = np.where(n > 0, key, np.iinfo(np.int64).max)
    order = np.argsort(key, axis=1, kind='stable')
    major, minor = order[:, 0], order[:, 1]
    minor = np.where(np.take_along_axis(n, minor[:, None], 1)[:, 0] > 0, minor, 255)
    lab = np.full(a.shape, UNCOVERED, dtype=np.uint8)
    covered = a != NOCOV
    lab[covered] = 0
    lab[covered & (a == major[:, None])] = 2
    lab[covered & (a == minor[:, None])] = 1
    return lab.reshape(shape)


def unit_to_mismatches(alleles_sr, kinds_s, positions=None, read_offset=0):
    """Raw alleles [S,R] of one unit -> the reference's ``mismatches`` dict."""
    S, R = alleles_sr.shape
    positions = positions if positions is not None else [1000 + 37 * s for s in range(S)]
    names = ["r%07d" % (read_offset + r) for r in range(R)]
    out = {}
    for s in range(S):
        row = alleles_sr[s]
        nt = {}
        for r in np.nonzero((row == ALT) | (row == THIRD))[0]:       # alts by first occurrence
            nt.setdefault(ALLELE_LETTER[int(row[r])], []).append(names[r])
        ref_reads = [names[r] for r in np.nonzero(row == REF)[0]]
        if ref_reads:
            nt[ALLELE_LETTER[REF]] = ref_reads                        # reference allele last
        out[positions[s]] = {
            'ref': ALLELE_LETTER[REF],
            'type': SITE_TYPE_NAMES[int(kinds_s[s])],
            'depth': {k: len(v) for k, v in nt.items()},
            'nt': nt,
            'neighbor': [], 'up': 'C', 'down': 'T',
        }
    return out


class SynthBatch:
    """G units of shape (S, R): raw alleles kept so that any unit can also be
    rendered in dict form."""

    def __init__(self, alleles, kinds):
        self.alleles, self.kinds = alleles, kinds

    @property
    def n_units(self):
        return self.alleles.shape[0]

    def mismatches(self, g):
        return unit_to_mismatches(self.alleles[g], self.kinds[g])

    def encoded(self, g) -> EncodedUnit:
        S = self.alleles.shape[1]
        return EncodedUnit([1000 + 37 * s for s in range(S)],
                           [SITE_TYPE_NAMES[int(k)] for k in self.kinds[g]],
                           labels_from_alleles(self.alleles[g]))

    def plane_batch(self) -> PlaneBatch:
        G, S, R = self.alleles.shape
        W = row_words(R)
        labels = labels_from_alleles(self.alleles)
        planes = pack_labels(labels.reshape(G * S, R)).reshape(-1)
        units = np.zeros(G, dtype=UNIT_DESC)
        units['plane_off'] = np.arange(G, dtype=np.uint64) * (3 * S * W)
        units['n_sites'], units['n_reads'], units['row_words'] = S, R, W
        units['site_off'] = np.arange(G, dtype=np.uint32) * S
        flags = self.kinds.reshape(-1).astype(np.uint8)
        flags = flags | np.where((labels.reshape(G * S, R) == 0).any(1), 4, 0).astype(np.uint8)
        return PlaneBatch(units, planes, flags)


def make_uniform(seed, G, S, R, cov, chunk=512) -> SynthBatch:
    """Deterministic batch of G same-shape units (cfg2 / cfg5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    al, kd = [], []
    for g0 in range(0, G, chunk):
        a, k = draw_alleles(rng, min(chunk, G - g0), S, R, cov)
        al.append(a)
        kd.append(k)
    return SynthBatch(np.concatenate(al), np.concatenate(kd))


def make_uniform_planes(seed, G, S, R, cov, chunk=512) -> PlaneBatch:
    """Same draw as make_uniform(...).plane_batch() without keeping the raw alleles."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W = row_words(R)
    planes = np.empty(G * 3 * S * W, dtype=np.uint32)
    flags = np.empty(G * S, dtype=np.uint8)
    for g0 in range(0, G, chunk):
        n = min(chunk, G - g0)
        a, k = draw_alleles(rng, n, S, R, cov)
        lab = labels_from_alleles(a).reshape(n * S, R)
        planes[g0 * 3 * S * W:(g0 + n) * 3 * S * W] = pack_labels(lab).reshape(-1)
        flags[g0 * S:(g0 + n) * S] = k.reshape(-1) | np.where((lab == 0).any(1), 4, 0).astype(np.uint8)
    units = np.zeros(G, dtype=UNIT_DESC)
    units['plane_off'] = np.arange(G, dtype=np.uint64) * (3 * S * W)
    units['n_sites'], units['n_reads'], units['row_words'] = S, R, W
    units['site_off'] = np.arange(G, dtype=np.uint32) * S
    return PlaneBatch(units, planes, flags)


def heavy_tail_shapes(rng, G, s_med=30, s_sigma=0.8, s_max=1000, r_med=150, r_sigma=1.0, r_max=20000):
    """cfg4 shapes: S ~ clip(lognormal(ln 30, .8), 2, 1000), R ~ clip(lognormal(ln 150, 1), 6, 20000)."""
    S = np.clip(np.round(rng.lognormal(np.log(s_med), s_sigma, G)), 2, s_max).astype(np.int64)
    R = np.clip(np.round(rng.lognormal(np.log(r_med), r_sigma, G)), 6, r_max).astype(np.int64)
    return S, R


def make_heavy_tail(seed, G, cov=0.5, keep_raw=False, **shape_kw):
    """Deterministic heavy-tailed batch (cfg4).  Returns (PlaneBatch, raw) where
    raw is a list of (alleles, kinds) per unit when keep_raw is set."""
    rng = np.random.Generator(np.random.PCG64(seed))
    S, R = heavy_tail_shapes(rng, G, **shape_kw)
    units = np.zeros(G, dtype=UNIT_DESC)
    chunks, flags, raw = [], [], []
    plane_off = site_off = 0
    for g in range(G):
        a, k = draw_alleles(rng, 1, int(S[g]), int(R[g]), cov)
        lab = labels_from_alleles(a[0])
        W = row_words(int(R[g]))
        units[g] = (plane_off, S[g], R[g], W, site_off)
        chunks.append(pack_labels(lab).reshape(-1))
        flags.append(site_flag_bytes([SITE_TYPE_NAMES[int(x)] for x in k[0]], lab))
        plane_off += 3 * int(S[g]) * W
        site_off += int(S[g])
        if keep_raw:
            raw.append((a[0], k[0]))
    pb = PlaneBatch(units, np.concatenate(chunks), np.concatenate(flags))
    return pb, raw


def make_deep_unit(seed, S, R, cov, keep_labels=False):
    """One deep unit (cfg3: 2 000 sites x 100 000 reads) from the same per-unit model.
    Returns (PlaneBatch, labels or None)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a, k = draw_alleles(rng, 1, S, R, cov)
    lab = labels_from_alleles(a[0])
    del a
    W = row_words(R)
    units = np.zeros(1, dtype=UNIT_DESC)
    units[0] = (0, S, R, W, 0)
    flags = site_flag_bytes([SITE_TYPE_NAMES[int(x)] for x in k[0]], lab)
    pb = PlaneBatch(units, pack_labels(lab).reshape(-1), flags)
    return pb, (lab if keep_labels else None)
