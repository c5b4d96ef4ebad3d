"""Import hot-path submodules of the reference pysteps WITHOUT running ``pysteps/__init__.py``
(needs jsmin/matplotlib, absent here) -- thin front of oracle/refimport.py, which finds the
reference at /root/reference (build container) or compiled in oracle/_ref (GPU box).  Used by the
``gen_*.py`` fixture generators in this directory and by tests that compare with the live
reference (skipped where neither exists)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import refimport as _r  # noqa: E402

REF = _r.SRC


def available():
    return _r.available()


def import_reference(vet_build_dir=None):
    return _r.import_reference(vet_build_dir)


def ref_module(name, vet_build_dir=None):
    return _r.ref_module(name, vet_build_dir)
