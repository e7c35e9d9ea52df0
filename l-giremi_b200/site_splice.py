"""Drop-in for the reference's `calculate_site_splice_mi` console script
(giremi/script/calculate_site_splice_mi.py:34-130): same two input tables, same output table
`<prefix>.site_splice_pair` (columns chromosome, site_pos, seq, splice_pos, count, mi), with the
mutual information of every (site allele, splice site) pair computed in ONE GPU submit
(api.site_splice_mutual_info) instead of a Python loop with `a in list` scans per pair (:106-125).

    python -m l-giremi_b200.site_splice -m READ_SITE.tsv -s READ_SPLICE.tsv -o PREFIX   (via importlib / runpy)

The table bookkeeping around the MI (which pairs exist, how many reads cover each) is host work
either way and follows the script's own procedure, including its row order: the splice table is
consumed in chunks of 10 000 rows, each chunk contributes its (chromosome, site_pos, seq,
splice_pos) groups in sorted order, a pair keeps the place of its first appearance."""
from __future__ import annotations

import argparse
from collections import defaultdict

import pandas as pd

from . import api

SPLICE_CHUNK_ROWS = 10000          # calculate_site_splice_mi.py:45-48


def parse_args(argv=None):
    parser = argparse.ArgumentParser(
        description='calculate the mutual information of mismatch site and splice site pairs')
    parser.add_argument("-m", "--read_site", type=str,
                        help="read-site file: [read_name, chromosome, pos, seq]")
    parser.add_argument("-s", "--read_splice", type=str,
                        help="corrected read-splice file: [read_name, chromosome, pos, type, corrected_pos, annotation]")
    parser.add_argument("-o", "--output_prefix", type=str, default='out', help="prefix of output file")
    return parser.parse_args(argv)


def read_tables(site_file, splice_file):
    """(pair rows [(chromosome, site_pos, seq, splice_pos, count)] in the script's order,
    sites {label: {allele: [read names]}}, splices {label: [read names]})  -- :40-102."""
    rsite = pd.read_table(site_file, header=0, sep='\t')
    count_of = {}                                   # insertion order = first appearance, as the script's Counter
    for chunk in pd.read_table(splice_file, header=0, sep='\t', chunksize=SPLICE_CHUNK_ROWS):
        merged = pd.merge(rsite, chunk[['read_name', 'chromosome', 'corrected_pos']], how='inner',
                          on=['read_name', 'chromosome'])
        if not len(merged):
            continue
        groups = merged.groupby(['chromosome', 'pos', 'seq', 'corrected_pos'])['read_name'].count()
        for key, n in zip(groups.index.tolist(), groups.tolist()):
            key = tuple(str(k) for k in key)        # the script joins the key as text and splits it again (:63-85)
            count_of[key] = count_of.get(key, 0) + int(n)
    sites = defaultdict(lambda: defaultdict(list))
    for chrom, pos, seq, name in zip(rsite['chromosome'].tolist(), rsite['pos'].tolist(), rsite['seq'].tolist(),
                                     rsite['read_name'].tolist()):
        sites['{}:{}'.format(chrom, pos)][seq].append(name)
    splices = defaultdict(list)
    with open(splice_file, 'r') as fh:              # every line, the header included (:97-102)
        for line in fh:
            cols = line.strip().split('\t')
            splices[':'.join([cols[1], cols[4]])].append(cols[0])
    rows = [key + (n,) for key, n in count_of.items()]
    return rows, sites, splices


def site_splice_table(site_file, splice_file, ctx=None) -> pd.DataFrame:
    """The script's output table, MI from the GPU."""
    rows, sites, splices = read_tables(site_file, splice_file)
    # alleles are looked up as the script does: by the text of the pair row (:107-116)
    by_text = {label: {str(seq): seq for seq in alleles} for label, alleles in sites.items()}
    triples = [(chrom + ':' + site_pos, by_text[chrom + ':' + site_pos][seq], chrom + ':' + splice_pos)
               for chrom, site_pos, seq, splice_pos, _n in rows]
    mi = api.site_splice_mutual_info(sites, splices, triples, ctx=ctx)
    out = pd.DataFrame([list(r[:4]) for r in rows], columns=['chromosome', 'site_pos', 'seq', 'splice_pos'])
    out.loc[:, 'count'] = [r[4] for r in rows]
    out.loc[:, 'mi'] = mi
    return out


def main(argv=None):
    args = parse_args(argv)
    table = site_splice_table(args.read_site, args.read_splice)
    table.to_csv(args.output_prefix + '.site_splice_pair', sep='\t', index=False)


if __name__ == '__main__':
    main()
