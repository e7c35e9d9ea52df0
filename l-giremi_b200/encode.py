"""Host-side encoder: the reference's ``mismatches[strand]`` dict -> bit-planes.

Replicates, before any GPU work, the dict semantics the kernels cannot see
(/root/reference/src/giremi/mutual_information.py):

  * :15-16  a read name listed twice at a site keeps its LAST allele in
            ('nt' key order, list order)
  * :25-32  major / minor allele by site-wide ``depth`` descending with the
            stable-sort tie-break (dict order); ``depth`` may count duplicates
  * :33-38  every other allele -> label 0 ("other")

Plane layout is documented in include/lgmi.h."""
from __future__ import annotations

import numpy as np

from ._lib import (SITE_HAS_OTHER, SITE_TYPE_CODE, SITE_TYPE_NAMES, UNIT_DESC)

UNCOVERED = 255  # label value for "read does not cover the site"


def row_words(n_reads: int) -> int:
    """32-bit words per plane row: whole 16-byte vectors."""
    return 4 * ((int(n_reads) + 127) // 128)


class EncodedUnit:
    """One (footprint, strand) unit in label-matrix form."""

    __slots__ = ("positions", "types", "labels", "bad_sites")

    def __init__(self, positions, types, labels, bad_sites=()):
        self.positions = list(positions)           # ascending
        self.types = list(types)                   # 'mismatch' | 'snp' | 'het_snp'
        self.labels = labels                       # uint8 [S, R]: 0/1/2 or UNCOVERED
        self.bad_sites = frozenset(bad_sites)      # indices whose 'depth' has < 2 alleles

    @property
    def n_sites(self):
        return self.labels.shape[0]

    @property
    def n_reads(self):
        return self.labels.shape[1]


def encode_mismatches(mismatches) -> EncodedUnit:
    """dict -> EncodedUnit.  Never mutates the input."""
    positions = sorted(mismatches)
    read_index = {}
    per_site = []
    for pos in positions:
        site = mismatches[pos]
        allele_of = {}
        for allele, names in site['nt'].items():       # last occurrence wins
            for name in names:
                allele_of[name] = allele
        for name in allele_of:
            if name not in read_index:
                read_index[name] = len(read_index)
        per_site.append(allele_of)
    labels = np.full((len(positions), len(read_index)), UNCOVERED, dtype=np.uint8)
    types, bad = [], []
    for s, pos in enumerate(positions):
        site = mismatches[pos]
        types.append(site['type'])
        ranked = sorted(site['depth'].items(), key=lambda kv: -kv[1])   # stable
        major = ranked[0][0] if len(ranked) > 0 else None
        minor = ranked[1][0] if len(ranked) > 1 else None
        if len(ranked) < 2:
            bad.append(s)      # the reference raises IndexError at :30 when a
                               # surviving pair touches this site
        row = labels[s]
        for name, allele in per_site[s].items():
            row[read_index[name]] = 2 if allele == major else (1 if allele == minor else 0)
    return EncodedUnit(positions, types, labels, bad)


def pack_labels(labels: np.ndarray) -> np.ndarray:
    """uint8 [S, R] label matrix -> uint32 [S, 3, W] planes (M, m, C)."""
    S, R = labels.shape
    W = row_words(R)
    bits = np.zeros((S, 3, W * 32), dtype=np.uint8)
    bits[:, 0, :R] = labels == 2
    bits[:, 1, :R] = labels == 1
    bits[:, 2, :R] = labels != UNCOVERED
    packed = np.packbits(bits, axis=-1, bitorder='little')
    return np.ascontiguousarray(packed).view('<u4').reshape(S, 3, W)


def site_flag_bytes(types, labels) -> np.ndarray:
    flags = np.array([SITE_TYPE_CODE[t] for t in types], dtype=np.uint8)
    if len(flags):
        flags |= np.where((labels == 0).any(axis=1), SITE_HAS_OTHER, 0).astype(np.uint8)
    return flags


class PlaneBatch:
    """Many units packed for one submit (the layout of include/lgmi.h)."""

    def __init__(self, units, planes, site_flags, positions=None, types=None, bad_sites=None):
        self.units = np.ascontiguousarray(units, dtype=UNIT_DESC)
        self.planes = np.ascontiguousarray(planes, dtype=np.uint32)
        self.site_flags = np.ascontiguousarray(site_flags, dtype=np.uint8)
        self.positions = positions      # list of per-unit position lists (optional metadata)
        self.types = types
        self.bad_sites = bad_sites

    @property
    def n_units(self):
        return len(self.units)

    @property
    def n_sites(self):
        return len(self.site_flags)

    @property
    def n_candidates(self):
        s = self.units['n_sites'].astype(np.int64)
        return int((s * (s - 1) // 2).sum())

    def packed2(self, tight=False) -> np.ndarray:
        """The planes in the two-plane form of lgmi_pipeline_step_packed: per site row [b0 | b1]
        with b0 = major | other, b1 = minor | other (2 bits per read: 00 not covered, 01 major,
        10 minor, 11 other).  Two thirds of the bytes of the [M | m | C] form; with tight=True
        (LGMI_MODE_TIGHT_INPUT) the rows are ceil(R/32) words wide instead of W = 4*ceil(R/128) and the
        units lie back to back: no 128-read padding on the wire."""
        S = self.units['n_sites'].astype(np.int64)
        W = self.units['row_words'].astype(np.int64)
        Wt = (self.units['n_reads'].astype(np.int64) + 31) // 32
        off = self.units['plane_off'].astype(np.int64)
        if tight:
            dst_off = np.concatenate(([0], np.cumsum(2 * S * Wt)))
            out = np.empty(int(dst_off[-1]), dtype=np.uint32)
        else:
            dst_off = off // 3 * 2
            out = np.empty(self.planes.size // 3 * 2, dtype=np.uint32)
        for w, wt in sorted(set(zip(W[S > 0].tolist(), Wt[S > 0].tolist()))):   # all units of one row width at once
            sel = np.flatnonzero((W == w) & (Wt == wt) & (S > 0))
            wo = wt if tight else w                                              # words per output plane row
            rows = np.concatenate([off[k] + 3 * w * np.arange(S[k]) for k in sel])
            drow = np.concatenate([dst_off[k] + 2 * wo * np.arange(S[k]) for k in sel])
            idx = rows[:, None] + np.arange(wo)[None, :]
            M, m, C = self.planes[idx], self.planes[idx + w], self.planes[idx + 2 * w]
            other = C & ~M & ~m
            didx = drow[:, None] + np.arange(wo)[None, :]
            out[didx] = M | other
            out[didx + wo] = m | other
        return out

    def site_types(self, unit):
        off = int(self.units['site_off'][unit])
        n = int(self.units['n_sites'][unit])
        return [SITE_TYPE_NAMES[f & 3] for f in self.site_flags[off:off + n]]

    def subset(self, index):
        """A new batch holding units[index] (planes re-packed contiguously)."""
        index = np.asarray(index)
        units = self.units[index].copy()
        chunks, flags = [], []
        plane_off = 0
        site_off = 0
        for k, u in enumerate(units):
            n = 3 * int(u['n_sites']) * int(u['row_words'])
            chunks.append(self.planes[int(u['plane_off']):int(u['plane_off']) + n])
            flags.append(self.site_flags[int(u['site_off']):int(u['site_off']) + int(u['n_sites'])])
            units[k]['plane_off'] = plane_off
            units[k]['site_off'] = site_off
            plane_off += n
            site_off += int(u['n_sites'])
        planes = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
        site_flags = np.concatenate(flags) if flags else np.zeros(0, np.uint8)

        def pick(meta):
            return [meta[i] for i in index] if meta is not None else None

        return PlaneBatch(units, planes, site_flags, pick(self.positions), pick(self.types),
                          pick(self.bad_sites))


def pack_units(encoded_units) -> PlaneBatch:
    """List[EncodedUnit] -> PlaneBatch."""
    units = np.zeros(len(encoded_units), dtype=UNIT_DESC)
    chunks, flags = [], []
    plane_off = 0
    site_off = 0
    for k, eu in enumerate(encoded_units):
        S, R = eu.labels.shape
        W = row_words(R)
        units[k] = (plane_off, S, R, W, site_off)
        if S:
            chunks.append(pack_labels(eu.labels).reshape(-1))
            flags.append(site_flag_bytes(eu.types, eu.labels))
        plane_off += 3 * S * W
        site_off += S
    planes = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
    site_flags = np.concatenate(flags) if flags else np.zeros(0, np.uint8)
    return PlaneBatch(units, planes, site_flags,
                      [eu.positions for eu in encoded_units],
                      [eu.types for eu in encoded_units],
                      [eu.bad_sites for eu in encoded_units])


def encode_batch(list_of_mismatches) -> PlaneBatch:
    return pack_units([encode_mismatches(m) for m in list_of_mismatches])


# --------------------------------------------------------------------------- #
# native host pieces (csrc/lgmi_host.inl): no device needed
# --------------------------------------------------------------------------- #
def cs_read_mismatches(cs_tag, reference_start, min_dist_from_splice=4):
    """One read's cs tag -> (mismatches, introns): ``[[pos, 'AG'], ...]`` after the
    splice-distance filter of mismatch.py:99-141 (``'AG'`` = reference base, read base, upper
    case, as stored at :143-147) and ``[[start, end], ...]`` of its introns in contig
    coordinates (cs.py:603-607).  Raises ValueError for a mark the reference's parser does
    not know (it raises KeyError there)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    raw = cs_tag.encode() if isinstance(cs_tag, str) else bytes(cs_tag)
    cap = max(1, raw.count(b"*"))
    cap_i = max(1, raw.count(b"~"))
    pos = np.empty(cap, np.int64)
    ref = np.empty(cap, "S1")
    alt = np.empty(cap, "S1")
    lo = np.empty(cap_i, np.int64)
    hi = np.empty(cap_i, np.int64)
    n, ni = C.c_uint32(), C.c_uint32()
    rc = lib.lgmi_cs_scan(raw, len(raw), int(reference_start), int(min_dist_from_splice), cap, _lib.ptr(pos),
                          _lib.ptr(ref), _lib.ptr(alt), C.byref(n), cap_i, _lib.ptr(lo), _lib.ptr(hi), C.byref(ni))
    if rc != 0:
        raise ValueError("cs tag not understood (liblgmi error %d): %r" % (rc, cs_tag[:60]))
    mm = [[int(p), (r + a).decode()] for p, r, a in zip(pos[:n.value].tolist(), ref[:n.value].tolist(),
                                                         alt[:n.value].tolist())]
    return mm, [[int(a), int(b)] for a, b in zip(lo[:ni.value].tolist(), hi[:ni.value].tolist())]


def encode_mismatches_native(mismatches):
    """dict -> (PlaneBatch of one unit, positions, types, bad_sites) through lgmi_encode_unit:
    the same planes and flags as pack_units([encode_mismatches(m)]), without the label matrix."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    positions = sorted(mismatches)
    S = len(positions)
    site_type = np.empty(S, np.uint8)
    n_depth = np.empty(S, np.uint32)
    n_nt = np.empty(S, np.uint32)
    d_allele, d_value, nt_allele, nt_count, names, types = [], [], [], [], [], []
    for s, pos in enumerate(positions):
        site = mismatches[pos]
        types.append(site['type'])
        site_type[s] = SITE_TYPE_CODE[site['type']]
        ids = {}
        n_depth[s] = len(site['depth'])
        for allele, depth in site['depth'].items():
            d_allele.append(ids.setdefault(allele, len(ids)))
            d_value.append(int(depth))
        n_nt[s] = len(site['nt'])
        for allele, lst in site['nt'].items():
            nt_allele.append(ids.setdefault(allele, len(ids)))
            nt_count.append(len(lst))
            names.extend(lst)
    blob = "\n".join(names).encode()
    d_allele = np.asarray(d_allele, np.uint32)
    d_value = np.asarray(d_value, np.int64)
    nt_allele = np.asarray(nt_allele, np.uint32)
    nt_count = np.asarray(nt_count, np.uint32)
    cap = 3 * S * row_words(len(names))
    planes = np.empty(max(1, cap), np.uint32)
    flags = np.empty(max(1, S), np.uint8)
    bad = np.empty(max(1, S), np.uint8)
    R, W = C.c_uint32(), C.c_uint32()
    rc = lib.lgmi_encode_unit(S, _lib.ptr(site_type), _lib.ptr(n_depth), _lib.ptr(d_allele), _lib.ptr(d_value),
                              _lib.ptr(n_nt), _lib.ptr(nt_allele), _lib.ptr(nt_count), blob, len(blob), cap,
                              _lib.ptr(planes), _lib.ptr(flags), _lib.ptr(bad), C.byref(R), C.byref(W))
    if rc != 0:
        raise RuntimeError("lgmi_encode_unit failed (%d)" % rc)
    units = np.zeros(1, dtype=UNIT_DESC)
    units[0] = (0, S, R.value, W.value, 0)
    pb = PlaneBatch(units, planes[:3 * S * W.value].copy(), flags[:S].copy(), [positions], [types],
                    [frozenset(np.flatnonzero(bad[:S]).tolist())])
    return pb


def concat_plane_batches(batches) -> PlaneBatch:
    """Several PlaneBatches -> one, units back to back in the given order (what the pipelined
    step wants)."""
    units, planes, flags, pos, typ, bad = [], [], [], [], [], []
    plane_off = site_off = 0
    for pb in batches:
        u = pb.units.copy()
        u['plane_off'] += plane_off - (int(pb.units['plane_off'][0]) if len(pb.units) else 0)
        u['site_off'] += site_off - (int(pb.units['site_off'][0]) if len(pb.units) else 0)
        units.append(u)
        planes.append(pb.planes)
        flags.append(pb.site_flags)
        plane_off += pb.planes.size
        site_off += pb.site_flags.size
        n = len(pb.units)
        pos += pb.positions if pb.positions is not None else [None] * n
        typ += pb.types if pb.types is not None else [None] * n
        bad += pb.bad_sites if pb.bad_sites is not None else [frozenset()] * n
    return PlaneBatch(np.concatenate(units) if units else np.zeros(0, UNIT_DESC),
                      np.concatenate(planes) if planes else np.zeros(0, np.uint32),
                      np.concatenate(flags) if flags else np.zeros(0, np.uint8), pos, typ, bad)
