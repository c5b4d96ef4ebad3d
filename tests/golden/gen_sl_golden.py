"""Generate tests/golden/sl_golden.npz by running the REFERENCE
(pysteps/extrapolation/semilagrangian.py, scipy 1.18.1) in this container.

    python tests/golden/gen_sl_golden.py

Inputs are rebuilt from seeds by tests (pysteps_b200._synthetic), only the
reference outputs are stored.  /root/reference is needed to run this script and
is NOT needed to run the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from sl_cases import CASES, SPLINE_CASES, build_case  # noqa: E402


def main():
    ref = ref_module("pysteps.extrapolation.semilagrangian")
    out = {}
    for name in CASES + SPLINE_CASES:
        args, kwargs = build_case(name)
        res = ref.extrapolate(*args, **kwargs)
        if isinstance(res, tuple):
            if res[0] is not None:
                out[name + "/out"] = res[0]
            out[name + "/disp"] = res[1]
        else:
            out[name + "/out"] = res
    path = os.path.join(HERE, "sl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
