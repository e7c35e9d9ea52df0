#!/usr/bin/env python
"""Per-kernel device time of the LAST step in an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...): kernel, microseconds, share.

    python tools/launch_summary.py gpurun_out/launches.csv [...]"""
import collections
import csv
import sys


def summarize(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for row in csv.DictReader(lines):
        try:
            rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row["Metric Unit"]))
        except (KeyError, ValueError):
            pass
    starts = [k for k, r in enumerate(rows) if "k_run_init" in r[0]]
    print(path, "%d launches, %d steps" % (len(rows), len(starts)))
    if not starts:
        return
    agg, total = collections.OrderedDict(), 0.0
    for name, v, unit in rows[starts[-1]:]:
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = name.split("(")[0]
        if "cutlass" in name or "at::" in name:          # the library GEMM / torch fills of the bench's own probes
            continue
        agg[name] = agg.get(name, 0.0) + v
        total += v
    for name, v in agg.items():
        print("   %-62s %9.1f us %5.1f%%" % (name[:62], v, 100 * v / total))
    print("   total %.1f us" % total)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarize(p)
