"""Stand-in for the `pysam` module over a simulated dataset (tools/simdata.py).

There is no htslib / pysam wheel in this image and no network, so BASELINE.json
configs[0] (the `l-giremi` CLI on a single-chromosome dataset) runs the
UNMODIFIED reference CLI with this module first on sys.path.  Every "file" the
CLI opens (-b, --genome_fasta, --snp_bcf, --annotation_gtf) is the same pickle
of a simdata.Dataset; the classes below expose exactly the calls the reference
makes (listed in tools/simdata.py).  Test / benchmark input infrastructure only."""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simdata  # noqa: E402

_cache = {}


def _load(path):
    if path not in _cache:
        with open(path, "rb") as fh:
            _cache[path] = pickle.load(fh)
    return _cache[path]


class AlignmentFile(simdata.AlignmentFile):
    def __init__(self, path, mode='rb'):
        super().__init__(_load(path))


class FastaFile(simdata.FastaFile):
    def __init__(self, path):
        super().__init__(_load(path))


class VariantFile(simdata.VariantFile):
    def __init__(self, path):
        super().__init__(_load(path))


class TabixFile(simdata.TabixFile):
    def __init__(self, path):
        super().__init__(_load(path))


asGTF = simdata.asGTF
