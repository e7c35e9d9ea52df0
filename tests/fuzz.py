"""Random ``mismatches`` dicts that exercise the reference's quirks (SURVEY 8a
Q1-Q6): third/fourth alleles, depth ties, duplicate read names within and
across alleles, depth != list length, sparse coverage around min_common."""
from __future__ import annotations

import numpy as np

TYPES = ('mismatch', 'snp', 'het_snp')
LETTERS = 'ACGT'


def random_mismatches(rng, n_sites=None, n_reads=None, cov=None, p_dup=0.15, p_multi=0.35,
                      p_depth_noise=0.2):
    S = int(n_sites if n_sites is not None else rng.integers(2, 9))
    R = int(n_reads if n_reads is not None else rng.integers(4, 40))
    cov = float(cov if cov is not None else rng.uniform(0.3, 1.0))
    names = ['q%03d' % r for r in range(R)]
    positions = sorted(rng.choice(np.arange(100, 100 + 50 * S), size=S, replace=False).tolist())
    out = {}
    for pos in positions:
        n_alleles = 2 + (int(rng.integers(0, 3)) if rng.random() < p_multi else 0)
        alleles = list(rng.permutation(list(LETTERS))[:n_alleles])
        weights = rng.dirichlet(np.ones(n_alleles) * 1.5)
        nt = {}
        for r in range(R):
            if rng.random() > cov:
                continue
            a = alleles[int(rng.choice(n_alleles, p=weights))]
            nt.setdefault(a, []).append(names[r])
            if rng.random() < p_dup:                  # same name again, maybe under another allele
                b = alleles[int(rng.integers(0, n_alleles))]
                nt.setdefault(b, []).append(names[r])
        while len(nt) < 2:                            # the filters guarantee two alleles
            a = alleles[len(nt)] if alleles[len(nt)] not in nt else alleles[0]
            nt.setdefault(a, []).append(names[int(rng.integers(0, R))])
            if len(nt) < 2:
                for a in alleles:
                    if a not in nt:
                        nt[a] = [names[int(rng.integers(0, R))]]
                        break
        depth = {a: len(v) for a, v in nt.items()}
        if rng.random() < p_depth_noise:              # depth need not equal the list length
            k = list(depth)[int(rng.integers(0, len(depth)))]
            depth[k] = max(1, depth[k] + int(rng.integers(-2, 3)))
        if rng.random() < 0.3:                        # force a tie for the stable-sort rule
            ks = list(depth)
            depth[ks[-1]] = depth[ks[0]]
        out[int(pos)] = {'ref': alleles[0], 'type': TYPES[int(rng.integers(0, 3))],
                         'depth': depth, 'nt': nt, 'neighbor': [], 'up': 'A', 'down': 'C'}
    return out


def hexf(x):
    return float(x).hex()


def unhex(s):
    return float.fromhex(s)
