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


def install():
    """Replace the reference's bindings in every giremi module already imported.
    Returns the list of (module, name) actually patched."""
    from . import api
    done = []
    for mod_name, attr in _TARGETS:
        mod = sys.modules.get(mod_name)
        if mod is None or not hasattr(mod, attr):
            continue
        _saved.setdefault((mod_name, attr), getattr(mod, attr))
        setattr(mod, attr, getattr(api, attr))
        done.append((mod_name, attr))
    return done


def uninstall():
    for (mod_name, attr), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, fn)
        del _saved[(mod_name, attr)]
