"""Monkeypatch seam: bind the GPU-backed functions into a loaded `giremi`.

The reference imports the two MI functions *by name* into giremi.mismatch
(/root/reference/src/giremi/mismatch.py:7-8) and `ecdf` into the CLI module
(/root/reference/src/giremi/script/giremi.py:15), so those bindings -- not only
giremi.mutual_information -- have to be replaced.  CUDA is initialised lazily
on the first call, i.e. in whichever process calls (never before a fork)."""
from __future__ import annotations

import sys

_saved = {}

_TARGETS = (
    ("giremi.mutual_information", "mismatch_pair_mutual_info"),
    ("giremi.mutual_information", "mean_mismatch_pair_mutual_info"),
    ("giremi.mismatch", "mismatch_pair_mutual_info"),
    ("giremi.mismatch", "mean_mismatch_pair_mutual_info"),
    ("giremi.stat", "ecdf"),
    ("giremi.script.giremi", "ecdf"),
)


# batched seam (SURVEY 8b ii): the per-region analysis and the per-chunk loop of the CLI
_BATCHED_TARGETS = (
    ("giremi.mismatch", "region_mismatch_analysis"),
    ("giremi.script.giremi", "region_mismatch_analysis"),
    ("giremi.script.giremi", "footprint_bulk_calculation"),
    ("giremi.script.giremi", "main"),
)


def install(batched=False):
    """Replace the reference's bindings in every giremi module already imported.
    With batched=True also `region_mismatch_analysis` (mismatch.py:345), the CLI's
    `footprint_bulk_calculation` (giremi.py:20) -- each chunk of footprints one GPU submit --
    and the CLI's `main` (giremi.py:324): workers extract, the parent owns every GPU and runs
    the mip pass in one launch.  Callers that hold their own reference to the stock `main`
    keep it (it then runs the patched per-chunk function in its workers).  Returns the list of
    (module, name) actually patched."""
    from . import api, batched as batched_mod
    done = []
    todo = [(m, a, api) for m, a in _TARGETS]
    if batched:
        todo += [(m, a, batched_mod) for m, a in _BATCHED_TARGETS]
    for mod_name, attr, source in todo:
        mod = sys.modules.get(mod_name)
        if mod is None or not hasattr(mod, attr):
            continue
        _saved.setdefault((mod_name, attr), getattr(mod, attr))
        setattr(mod, attr, getattr(source, attr))
        done.append((mod_name, attr))
    return done


def uninstall():
    for (mod_name, attr), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, fn)
        del _saved[(mod_name, attr)]
