"""Synthetic single-chromosome long-read dataset in the shape of BASELINE.json
configs[0], plus in-memory stand-ins for the four pysam
objects the reference opens (AlignmentFile, FastaFile, VariantFile, TabixFile).

Test and benchmark INPUT infrastructure: it produces what the reference reads
(spliced reads with short-form cs tags, a genome, a GTF, a SNP list, a repeat
table); nothing here computes anything of the MI step.  The stand-ins implement
only the calls the reference makes:

  AlignmentFile.fetch(chrom[, start, end])  -> reads (query_name, is_reverse,
      reference_start, reference_end, get_tag('cs'))
      giremi/mismatch.py:66-90, strand.py:146-165, footprint.py:20-24
  AlignmentFile.pileup(contig=, start=, stop=) -> columns (pos,
      get_query_names(), get_query_sequences())        giremi/mismatch.py:160-188
  FastaFile.fetch(chrom, start, end)                    giremi/mismatch.py:284-291
  VariantFile.fetch(chrom, start, end) -> records with .start   fileio.py:24-30
  TabixFile.fetch(chrom, start, end, parser=asGTF()) -> entries with feature,
      gene_name, start, end, strand                     strand.py:181-193

Model: genes 50 kb apart on one contig, 3-6 exons with gt..ag introns, reads
covering most of the transcript, per-read haplotype, het SNPs that follow the
haplotype, A>G editing sites (T>C on the minus strand) with Beta(2,5) levels of
which 30 % are haplotype-linked, 0.2 % substitution noise.
"""
from __future__ import annotations

import bisect
from collections import defaultdict

import numpy as np

COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


class Read:
    __slots__ = ("query_name", "is_reverse", "reference_start", "reference_end", "cs", "pos", "seq")

    def __init__(self, name, is_reverse, start, end, cs, pos, seq):
        self.query_name, self.is_reverse = name, is_reverse
        self.reference_start, self.reference_end = start, end
        self.cs, self.pos, self.seq = cs, pos, seq      # pos: covered genome positions, seq: read bases there

    def get_tag(self, tag):
        if tag != 'cs':
            raise KeyError(tag)
        return self.cs


class PileupColumn:
    __slots__ = ("pos", "_names", "_seqs")

    def __init__(self, pos, names, seqs):
        self.pos, self._names, self._seqs = pos, names, seqs

    def get_query_names(self):
        return list(self._names)

    def get_query_sequences(self):
        return list(self._seqs)


class AlignmentFile:
    def __init__(self, dataset, mode='rb'):
        self.ds = dataset.ds if isinstance(dataset, AlignmentFile) else dataset
        self._reads = self.ds.reads
        self._starts = [r.reference_start for r in self._reads]
        self._max_len = max((r.reference_end - r.reference_start for r in self._reads), default=0)

    def fetch(self, contig=None, start=None, stop=None):
        if contig is not None and contig != self.ds.chrom:
            return
        if start is None:
            yield from self._reads
            return
        lo = bisect.bisect_left(self._starts, start - self._max_len)
        for r in self._reads[lo:]:
            if r.reference_start >= stop:
                break
            if r.reference_end > start:
                yield r

    def pileup(self, contig=None, start=None, stop=None):
        cols = defaultdict(lambda: ([], []))
        for r in self.fetch(contig, start, stop):
            for p, b in zip(r.pos, r.seq):
                c = cols[p]
                c[0].append(r.query_name)
                c[1].append(b)
        for p in sorted(cols):
            yield PileupColumn(p, cols[p][0], cols[p][1])

    def close(self):
        pass


class FastaFile:
    def __init__(self, dataset):
        self.ds = dataset

    def fetch(self, contig, start, end):
        return self.ds.genome[max(0, start):end]

    def close(self):
        pass


class _Variant:
    __slots__ = ("start",)

    def __init__(self, start):
        self.start = start


class VariantFile:
    def __init__(self, dataset):
        self.ds = dataset

    def fetch(self, contig, start=None, end=None):
        for p in self.ds.snp_positions:
            if start is None or start <= p < end:
                yield _Variant(p)

    def close(self):
        pass


class _GtfEntry:
    __slots__ = ("feature", "gene_name", "start", "end", "strand")

    def __init__(self, feature, gene_name, start, end, strand):
        self.feature, self.gene_name, self.start, self.end, self.strand = feature, gene_name, start, end, strand


class TabixFile:
    def __init__(self, dataset):
        self.ds = dataset

    def fetch(self, contig, start, end, parser=None):
        for e in self.ds.gtf:
            if e.end > start and e.start < end:
                yield e

    def close(self):
        pass


def asGTF():
    return None


class Dataset:
    """chrom, genome (str), reads (sorted by start), gtf entries, snp positions, repeats."""

    def __init__(self, seed=20261018, n_genes=8, reads_per_gene=120, chrom="chr1", gene_spacing=50000,
                 n_het=(3, 8), n_edit=(10, 40), noise=0.002):
        rng = np.random.Generator(np.random.PCG64(seed))
        self.chrom = chrom
        L = 20000 + gene_spacing * n_genes + 20000
        g = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)
        reads, gtf, snps = [], [], []
        truth = []
        for gi in range(n_genes):
            strand = '+' if rng.random() < 0.5 else '-'
            pos = 20000 + gene_spacing * gi
            n_ex = int(rng.integers(3, 7))
            exons = []
            for e in range(n_ex):
                elen = int(rng.integers(150, 400))
                exons.append((pos, pos + elen))
                pos += elen
                if e + 1 < n_ex:
                    ilen = int(rng.integers(300, 1500))
                    # canonical splice signals as seen on the forward genome
                    sig = (b"GT", b"AG") if strand == '+' else (b"CT", b"AC")
                    g[pos:pos + 2] = np.frombuffer(sig[0], dtype=np.uint8)
                    g[pos + ilen - 2:pos + ilen] = np.frombuffer(sig[1], dtype=np.uint8)
                    pos += ilen
            name = "gene%03d" % gi
            gtf.append(_GtfEntry('gene', name, exons[0][0], exons[-1][1], strand))
            for a, b in exons:
                gtf.append(_GtfEntry('exon', name, a, b, strand))
            tpos = np.concatenate([np.arange(a, b) for a, b in exons])         # transcript -> genome
            exon_of = np.concatenate([np.full(b - a, k) for k, (a, b) in enumerate(exons)])
            T = len(tpos)
            inner = np.array([t for t in range(T) if min(tpos[t] - exons[exon_of[t]][0],
                                                         exons[exon_of[t]][1] - 1 - tpos[t]) >= 8])
            # het SNPs
            k_het = int(rng.integers(n_het[0], n_het[1] + 1))
            het_t = rng.choice(inner, k_het, replace=False)
            het_alt = {}
            for t in het_t:
                ref = chr(g[tpos[t]])
                het_alt[int(t)] = str(rng.choice([b for b in "ACGT" if b != ref]))
                snps.append(int(tpos[t]))
            # editing sites: A on the transcript strand (T on the genome for '-')
            edit_base = ord('A') if strand == '+' else ord('T')
            edit_to = 'G' if strand == '+' else 'C'
            cand = [int(t) for t in inner if g[tpos[t]] == edit_base and int(t) not in het_alt]
            k_ed = min(len(cand), int(rng.integers(n_edit[0], n_edit[1] + 1)))
            ed_t = rng.choice(cand, k_ed, replace=False) if k_ed else []
            ed_level = {int(t): float(rng.beta(2, 5)) for t in ed_t}
            ed_linked = {int(t): bool(rng.random() < 0.3) for t in ed_t}
            truth.append(dict(gene=name, strand=strand, het=[int(tpos[t]) for t in het_t],
                              edit=[int(tpos[t]) for t in ed_t]))
            for ri in range(reads_per_gene):
                hap = int(rng.random() < 0.5)
                t0 = int(rng.integers(0, max(1, int(0.3 * T))))
                t1 = int(rng.integers(int(0.7 * T), T)) + 1
                bases = g[tpos[t0:t1]].copy()
                u = rng.random(t1 - t0)
                for t in range(t0, t1):
                    if t in het_alt:
                        carries = hap == 1
                        if rng.random() < 0.01:
                            carries = not carries
                        if carries:
                            bases[t - t0] = ord(het_alt[t])
                    elif t in ed_level:
                        if (not ed_linked[t] or hap == 1) and rng.random() < ed_level[t] * (2.0 if ed_linked[t] else 1.0):
                            bases[t - t0] = ord(edit_to)
                    elif u[t - t0] < noise:
                        ref = chr(g[tpos[t]])
                        bases[t - t0] = ord(str(rng.choice([b for b in "ACGT" if b != ref])))
                # cs tag (short form) along the genome
                cs, run = [], 0
                for t in range(t0, t1):
                    if t > t0 and exon_of[t] != exon_of[t - 1]:
                        if run:
                            cs.append(":%d" % run)
                            run = 0
                        a, b = exons[exon_of[t - 1]][1], exons[exon_of[t]][0]
                        intr = bytes(g[a:b]).decode().lower()
                        cs.append("~%s%d%s" % (intr[:2], b - a, intr[-2:]))
                    ref = g[tpos[t]]
                    if bases[t - t0] == ref:
                        run += 1
                    else:
                        if run:
                            cs.append(":%d" % run)
                            run = 0
                        cs.append("*%s%s" % (chr(ref).lower(), chr(bases[t - t0]).lower()))
                if run:
                    cs.append(":%d" % run)
                is_rev = (strand == '-') if rng.random() < 0.9 else (strand == '+')
                reads.append(Read("%s_r%04d" % (name, ri), is_rev, int(tpos[t0]), int(tpos[t1 - 1]) + 1, "".join(cs),
                                  [int(p) for p in tpos[t0:t1]], [chr(b) for b in bases]))
        reads.sort(key=lambda r: (r.reference_start, r.query_name))
        self.genome = bytes(g).decode()
        self.reads, self.gtf = reads, gtf
        self.snp_positions = sorted(set(snps))
        self.truth = truth
        # a few simple-repeat intervals (3-column table, fileio.py:11-17)
        self.repeats = [[int(20000 + gene_spacing * k + 100), int(20000 + gene_spacing * k + 130)] for k in range(n_genes)]

    # the four objects footprint_bulk_calculation opens (giremi.py:21-24)
    def sam(self):
        return AlignmentFile(self)

    def fasta(self):
        return FastaFile(self)

    def vcf(self):
        return VariantFile(self)

    def tabix(self):
        return TabixFile(self)

    def footprints(self, min_read_count=2):
        """[chrom, start, end, n_reads] like giremi.footprint.get_footprints (footprint.py:20-29)."""
        out, cur = [], None
        for r in self.reads:
            if cur is None or cur[1] < r.reference_start:
                if cur is not None:
                    out.append(cur)
                cur = [r.reference_start, r.reference_end, 1]
            else:
                cur[1] = max(cur[1], r.reference_end)
                cur[2] += 1
        if cur is not None:
            out.append(cur)
        return [[self.chrom, a, b, n] for a, b, n in out if n >= min_read_count]

    def write_repeat_file(self, path):
        with open(path, "w") as fh:
            for a, b in self.repeats:
                fh.write("%s\t%d\t%d\n" % (self.chrom, a, b))
