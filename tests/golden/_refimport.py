"""Import hot-path submodules of the reference pysteps WITHOUT running
``pysteps/__init__.py`` (needs jsmin/matplotlib, absent here).  Only used by the
``gen_*.py`` fixture generators in this directory and by tests that are skipped
when ``/root/reference`` does not exist (it does not on the GPU box)."""
import importlib
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "pysteps"))


def import_reference(vet_build_dir=None):
    if "pysteps" in sys.modules and getattr(sys.modules["pysteps"], "__b200_stub__", False):
        return sys.modules["pysteps"]
    pk = types.ModuleType("pysteps")
    pk.__path__ = [os.path.join(REF, "pysteps")]
    pk.__b200_stub__ = True
    sys.modules["pysteps"] = pk
    mm = types.ModuleType("pysteps.motion")
    mm.__path__ = [os.path.join(REF, "pysteps", "motion")]
    if vet_build_dir:
        mm.__path__.append(os.path.join(vet_build_dir, "pysteps", "motion"))
    sys.modules["pysteps.motion"] = mm
    pk.motion = mm
    return pk


def ref_module(name, vet_build_dir=None):
    import_reference(vet_build_dir)
    return importlib.import_module(name)
