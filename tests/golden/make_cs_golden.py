#!/usr/bin/env python
"""Golden vectors for lgmi_cs_scan from the REAL reference (run in the build container, where
/root/reference exists): random short-form cs tags -> CS.get_mismatches / get_introns + FILTER 1
of mismatch.py:99-141, exactly as get_region_mismatches_with_filters applies them per read."""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"
pkg = types.ModuleType("giremi")
pkg.__path__ = [os.path.join(REF, "giremi")]
pkg.__version__ = "0.2.4"
sys.modules["giremi"] = pkg
from giremi.cs import CS  # noqa: E402
from giremi.utils import merge_intervals, positions_in_intervals  # noqa: E402


def reference(cs_tag, start, min_dist):
    read_CS = CS.from_cs_tag_string(cs_tag, "chr1", start, "+")
    read_mismatches = [[a[0], a[3]] for a in read_CS.get_mismatches(coordinate='contig')]
    read_mismatches.sort(key=lambda a: a[0])
    read_introns = read_CS.get_introns(coordinate='contig')
    read_introns.sort(key=lambda a: a[0])
    if len(read_mismatches) > 0 and len(read_introns) > 0 and min_dist > 0:
        splicing_pos = sorted([a[0] for a in read_introns] + [a[1] for a in read_introns])
        intervals, _ = merge_intervals([[a - min_dist, a + min_dist] for a in splicing_pos])
        inside, _ = positions_in_intervals([a[0] for a in read_mismatches], intervals)
    else:
        inside = [False for _ in read_mismatches]
    kept = [[int(p), v[0].upper() + v[1].upper()] for (p, v), i in zip(read_mismatches, inside) if not i]
    return kept, [[int(a[0]), int(a[1])] for a in read_introns]


def random_cs(rng):
    parts = []
    for _ in range(int(rng.integers(1, 40))):
        k = rng.random()
        if k < 0.45:
            parts.append(":%d" % rng.integers(1, 60))
        elif k < 0.75:
            a, b = rng.choice(list("acgtn"), 2, replace=False)
            parts.append("*%s%s" % (a, b))
        elif k < 0.82:
            parts.append("+" + "".join(rng.choice(list("acgt"), int(rng.integers(1, 6)))))
        elif k < 0.89:
            parts.append("-" + "".join(rng.choice(list("acgt"), int(rng.integers(1, 6)))))
        else:
            parts.append("~%s%d%s" % (rng.choice(["gt", "ct", "gc", "at"]), rng.integers(2, 40), rng.choice(["ag", "ac"])))
    return "".join(parts)


def main():
    rng = np.random.default_rng(20261101)
    cases = []
    for _ in range(400):
        cs = random_cs(rng)
        start = int(rng.integers(0, 10 ** 6))
        d = int(rng.choice([0, 1, 4, 10]))
        mm, introns = reference(cs, start, d)
        cases.append({"cs": cs, "start": start, "min_dist": d, "mismatches": mm, "introns": introns})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cs_scan.json")
    with open(out, "w") as fh:
        json.dump({"generator": "tests/golden/make_cs_golden.py (reference giremi 0.2.4)", "data": cases}, fh)
    print(out, len(cases))


if __name__ == "__main__":
    main()
