"""TEST INFRASTRUCTURE: host builds of CUDA kernel bodies (see spline_host.cpp)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhost_kernels.so")
        csrc = os.path.join(_HERE, "..", "..", "pysteps_b200", "csrc")
        units = [os.path.join(_HERE, "spline_host.cpp"), os.path.join(_HERE, "proesmans_host.cpp"),
                 os.path.join(_HERE, "knn_host.cpp")]
        srcs = units + [os.path.join(csrc, "spline_body.cuh"), os.path.join(csrc, "proesmans_body.cuh"),
                        os.path.join(csrc, "knn_body.cuh"), os.path.join(csrc, "quantise_body.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            tmp = f"{so}.tmp.{os.getpid()}"
            subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off",
                                   "-fno-fast-math", "-Wall", "-o", tmp] + units)
            os.replace(tmp, so)
        _LIB = ctypes.CDLL(so)
    return _LIB
