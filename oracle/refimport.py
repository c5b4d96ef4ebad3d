"""TEST INFRASTRUCTURE ONLY -- import submodules of the REFERENCE pysteps without running
``pysteps/__init__.py`` (it needs jsmin / matplotlib, absent in this image): a stub namespace
package ``pysteps`` whose path is the reference tree (``/root/reference/pysteps`` in the build
container) or its compiled form (``oracle/_ref/pysteps``, built by oracle/build_ref.py, which is
what exists on the GPU box).  Every submodule then runs its own, unmodified code."""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
BUILT = os.path.join(HERE, "_ref")


def roots():
    """package directories to search, source tree first"""
    r = []
    if os.path.isdir(os.path.join(SRC, "pysteps")):
        r.append(os.path.join(SRC, "pysteps"))
    if os.path.isdir(os.path.join(BUILT, "pysteps")):
        r.append(os.path.join(BUILT, "pysteps"))
    return r


def available(extensions=False):
    if extensions:
        mot = os.path.join(BUILT, "pysteps", "motion")
        return bool(roots()) and os.path.isdir(mot) and any(f.startswith("_vet.") and f.endswith(".so")
                                                            for f in os.listdir(mot))
    return bool(roots())


def import_reference(extra_motion_dir=None):
    if "pysteps" in sys.modules and getattr(sys.modules["pysteps"], "__b200_stub__", False):
        return sys.modules["pysteps"]
    rs = roots()
    if not rs:
        raise ImportError("the reference is neither at /root/reference nor built into oracle/_ref")
    pk = types.ModuleType("pysteps")
    pk.__path__ = list(rs)
    pk.__b200_stub__ = True
    sys.modules["pysteps"] = pk
    mm = types.ModuleType("pysteps.motion")
    mm.__path__ = [os.path.join(r, "motion") for r in rs]
    if extra_motion_dir:
        mm.__path__.append(os.path.join(extra_motion_dir, "pysteps", "motion"))
    sys.modules["pysteps.motion"] = mm
    pk.motion = mm
    return pk


def ref_module(name, extra_motion_dir=None):
    import_reference(extra_motion_dir)
    return importlib.import_module(name)
