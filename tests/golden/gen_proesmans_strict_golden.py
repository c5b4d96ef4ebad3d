"""Generate tests/golden/proesmans_strict_golden.npz: outputs of the reference SOURCE
(pysteps/motion/proesmans.py + _proesmans.pyx) with the extension compiled WITHOUT -ffast-math
(-O2 -fno-fast-math -ffp-contract=off), i.e. with defined IEEE rounding.  The oracle, with the
mean inconsistency accumulated in the source's raster order, must reproduce these bit for bit --
including the ill-conditioned cases in which the shipped -ffast-math build drifts by pixels.

    B=/tmp/proes_strict; mkdir -p $B/pysteps/motion; cp /root/reference/pysteps/motion/_proesmans.pyx $B/pysteps/motion/
    (setup: Extension("pysteps.motion._proesmans", ..., extra_compile_args=["-O2","-fno-fast-math","-ffp-contract=off"]))
    cd $B && CC=/usr/bin/gcc LDSHARED="/usr/bin/gcc -shared" python setup.py build_ext --inplace
    python tests/golden/gen_proesmans_strict_golden.py /tmp/proes_strict
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from proesmans_cases import STRICT_CASES, build_case  # noqa: E402


def main():
    build = sys.argv[1] if len(sys.argv) > 1 else "/tmp/proes_strict"
    ref = ref_module("pysteps.motion.proesmans", build)
    out = {}
    for name in STRICT_CASES:
        frames, kw = build_case(name)
        adv, q = ref.proesmans(frames[-2:], full_output=True, **kw)
        out[name + "/advfield"] = adv
        out[name + "/quality"] = q
    path = os.path.join(HERE, "proesmans_strict_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
