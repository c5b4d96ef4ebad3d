"""CPU oracle for the pysteps advection hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  The product package ``pysteps_b200`` never does; it
fails loudly when its CUDA library is missing instead of falling back here.

``oracle.lib()`` returns the ctypes handle of ``liboracle.so`` (plain C,
strict IEEE float64), building it with ``make`` on first use.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
    )
    if force or stale:
        # same flags as the Makefile; built under a private name and renamed, so that two
        # processes (e.g. a test run next to a bench run) never load a half-written library
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        tmp = f"{so}.tmp.{os.getpid()}"
        subprocess.check_call([cc, "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-Wall",
                               "-Wextra", "-std=c11", "-shared", "-o", tmp] + sorted(srcs) + ["-lm"])
        os.replace(tmp, so)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ora_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().ora_num_threads())


def set_num_threads(n):
    """OpenMP threads of the C oracle (all host cores: os.cpu_count())."""
    lib().ora_set_num_threads(int(n))
    return num_threads()
