#!/usr/bin/env python
"""Golden vectors for the site x splice-site MI from the REFERENCE ITSELF:
runs /root/reference/src/giremi/script/calculate_site_splice_mi.py main() (unmodified, loaded by
path; it imports only pandas + scikit-learn) on two small TSVs written here, and stores the TSV
rows together with the pairs table it writes (MI as hex floats) in tests/golden/site_splice.json.

    python tests/golden/make_site_splice_golden.py        # only in the container that has /root/reference

The inputs are committed inside the JSON, so the tests need neither the reference nor this script.

One environment shim, none in the reference: the script initialises its `mi` column with the integer -1
and then stores floats into it (:104, :123), which pandas < 3 silently upcasts and pandas 3 (installed
here) refuses with a TypeError.  For the duration of main() the pandas-2 behaviour is restored
(Block.coerce_to_target_dtype with raise_on_upcast=False); the values stored are the reference's own
mutual_info_score results either way."""
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/src/giremi/script/calculate_site_splice_mi.py"


def make_inputs(seed=20261040):
    rng = np.random.default_rng(seed)
    reads = ["read%04d" % k for k in range(260)]
    site_rows, splice_rows = [], []
    for chrom, n_sites, n_splices in (("chr1", 9, 5), ("chr2", 4, 3)):
        splice_pos = [50000 + 1013 * s for s in range(n_splices)]
        has = {}                                                   # (read, splice index) -> listed
        for r in reads:
            if rng.random() < 0.15:
                continue                                           # a read without splice records
            for si, sp in enumerate(splice_pos):
                if si == n_splices - 1 or rng.random() < 0.55:     # the last splice site: every spliced read has it
                    # [read_name, chromosome, pos, type, corrected_pos, annotation]
                    splice_rows.append([r, chrom, sp + int(rng.integers(-3, 4)), "intron_start", sp, "annotated"])
                    has[(r, si)] = True
        for k in range(n_sites):
            pos = 1000 + 371 * k
            p = rng.dirichlet([6, 3, 1, 0.3])
            for r in reads:
                if rng.random() < 0.6:
                    if k % 4 == 1:                                 # allele linked to the first splice site (95 %)
                        linked = (r, 0) in has
                        seq = ("G" if linked else "A") if rng.random() < 0.95 else ("A" if linked else "G")
                    elif k % 4 == 2:                               # ... exactly
                        seq = "C" if (r, 1) in has else "T"
                    else:
                        seq = str(rng.choice(list("ACGT"), p=p))
                    site_rows.append([r, chrom, pos, seq])
            if k % 3 == 0:                                         # the same read listed twice at a site
                site_rows.append([site_rows[-1][0], chrom, pos, "A"])
    return site_rows, splice_rows


def main():
    spec = importlib.util.spec_from_file_location("ref_site_splice", SCRIPT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    site_rows, splice_rows = make_inputs()
    with tempfile.TemporaryDirectory() as tmp:
        sf, pf, out = os.path.join(tmp, "site.tsv"), os.path.join(tmp, "splice.tsv"), os.path.join(tmp, "out")
        pd.DataFrame(site_rows, columns=["read_name", "chromosome", "pos", "seq"]).to_csv(sf, sep="\t", index=False)
        pd.DataFrame(splice_rows, columns=["read_name", "chromosome", "pos", "type", "corrected_pos", "annotation"]
                     ).to_csv(pf, sep="\t", index=False)
        argv = sys.argv
        sys.argv = ["calculate_site_splice_mi", "-m", sf, "-s", pf, "-o", out]
        from pandas.core.internals import blocks
        strict = blocks.Block.coerce_to_target_dtype
        blocks.Block.coerce_to_target_dtype = lambda self, other, raise_on_upcast=False: strict(self, other, False)
        try:
            mod.main()
        finally:
            sys.argv = argv
            blocks.Block.coerce_to_target_dtype = strict
        table = pd.read_table(out + ".site_splice_pair", dtype={"chromosome": str, "seq": str}, float_precision="round_trip")
    pairs = [[str(r.chromosome), int(r.site_pos), str(r.seq), int(r.splice_pos), int(r["count"]), float(r.mi).hex()]
             for _, r in table.iterrows()]
    doc = {"source": "calculate_site_splice_mi.py main() of gxiaolab/L-GIREMI v0.2.4, scikit-learn %s, pandas %s"
                     % (__import__("sklearn").__version__, pd.__version__),
           "data": {"site_rows": site_rows, "splice_rows": splice_rows,
                    "pair_columns": ["chromosome", "site_pos", "seq", "splice_pos", "count", "mi"], "pairs": pairs}}
    with open(os.path.join(HERE, "site_splice.json"), "w") as fh:
        json.dump(doc, fh)
    mi = np.array([float.fromhex(p[5]) for p in pairs])
    print("wrote %d pairs (%d with MI > 0, max %.4f) from %d site rows, %d splice rows"
          % (len(pairs), int((mi > 0).sum()), mi.max(), len(site_rows), len(splice_rows)))


if __name__ == "__main__":
    main()
