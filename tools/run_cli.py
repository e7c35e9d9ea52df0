#!/usr/bin/env python
"""Runs the reference's `l-giremi` CLI (giremi.script.giremi:main, unmodified, from
baseline/_ref) on a simulated dataset -- stock, or with this repository's GPU MI step
patched in (`--patched`: lg.install(batched=True) before main()).

    python tools/run_cli.py [--patched] DATASET.pkl OUT_PREFIX [l-giremi options ...]

Prints one JSON line with the wall time of main() and of the MI step inside it."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools", "pysam_shim"), os.path.join(ROOT, "baseline", "_ref"), ROOT]


def main():
    argv = sys.argv[1:]
    patched = "--patched" in argv
    argv = [a for a in argv if a != "--patched"]
    dataset, prefix, extra = argv[0], argv[1], argv[2:]
    import giremi.script.giremi as cli
    if patched:
        lg = importlib.import_module("l-giremi_b200")
        lg.install(batched=True)        # before the pool forks; CUDA itself starts lazily in whoever computes
    repeat = prefix + ".repeats.txt"
    import pickle
    with open(dataset, "rb") as fh:
        ds = pickle.load(fh)
    ds.write_repeat_file(repeat)
    sys.argv = ["l-giremi", "-b", dataset, "-c", ds.chrom, "-o", prefix, "--genome_fasta", dataset,
                "--snp_bcf", dataset, "--annotation_gtf", dataset, "--repeat_txt", repeat] + extra
    t0 = time.perf_counter()
    cli.main()
    wall = time.perf_counter() - t0
    print(json.dumps({"patched": patched, "wall_s": wall, "argv": extra}))


if __name__ == "__main__":
    main()
