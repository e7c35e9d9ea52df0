#!/usr/bin/env python
"""Runs the reference's `l-giremi` CLI (giremi.script.giremi, unmodified, from baseline/_ref) on a
simulated dataset -- stock, or with this repository's GPU MI step patched in (`--patched`:
lg.install(batched=True) before main(); `--patched-functions`: only the name-level drop-ins, the
stock main and its workers keep running).

    python tools/run_cli.py [--patched | --patched-functions] DATASET.pkl OUT_PREFIX [l-giremi options ...]

Prints one JSON line: wall time of main() and the time of the MI step inside it:
  stock     the two MI functions (mutual_information.py:6-60) wrapped with timers inside the pool
            workers: `mi_step_cpu_s` = seconds summed over all workers, `mi_calls` = units
  patched   `mi_step_s` = wall time of the parent's GPU step over all units (encode excluded,
            frames included), `mip_s` = the one-launch mip pass, `extract_pool_s` = the workers"""
import importlib
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools", "pysam_shim"), os.path.join(ROOT, "baseline", "_ref"), ROOT]

_mi_seconds = mp.Value('d', 0.0)      # created before the pool forks: shared with the workers
_mi_calls = mp.Value('l', 0)


def _timed(fn, count):
    def wrapper(*a, **kw):
        t0 = time.perf_counter()
        try:
            return fn(*a, **kw)
        finally:
            dt = time.perf_counter() - t0
            with _mi_seconds.get_lock():
                _mi_seconds.value += dt
            if count:
                with _mi_calls.get_lock():
                    _mi_calls.value += 1
    return wrapper


def main():
    argv = sys.argv[1:]
    patched = "--patched" in argv
    functions_only = "--patched-functions" in argv
    argv = [a for a in argv if a not in ("--patched", "--patched-functions")]
    dataset, prefix, extra = argv[0], argv[1], argv[2:]
    import giremi.mismatch as gm
    import giremi.script.giremi as cli
    lg = None
    if patched or functions_only:
        lg = importlib.import_module("l-giremi_b200")
        lg.install(batched=patched)     # before the pool forks; CUDA itself starts lazily in whoever computes
    if not patched:                     # stock / name-level drop-ins: time the two MI functions where they run
        gm.mismatch_pair_mutual_info = _timed(gm.mismatch_pair_mutual_info, True)
        gm.mean_mismatch_pair_mutual_info = _timed(gm.mean_mismatch_pair_mutual_info, False)
    repeat = prefix + ".repeats.txt"
    import pickle
    with open(dataset, "rb") as fh:
        ds = pickle.load(fh)
    ds.write_repeat_file(repeat)
    sys.argv = ["l-giremi", "-b", dataset, "-c", ds.chrom, "-o", prefix, "--genome_fasta", dataset,
                "--snp_bcf", dataset, "--annotation_gtf", dataset, "--repeat_txt", repeat] + extra
    t0 = time.perf_counter()
    cli.main()                          # (the attribute: install(batched=True) rebinds it)
    wall = time.perf_counter() - t0
    line = {"patched": patched, "patched_functions_only": functions_only, "wall_s": wall, "argv": extra}
    if patched:
        batched = importlib.import_module("l-giremi_b200.batched")
        line.update(batched.last_run_times)
    else:
        line.update({"mi_step_cpu_s": _mi_seconds.value, "mi_calls": int(_mi_calls.value)})
    print(json.dumps(line))


if __name__ == "__main__":
    main()
