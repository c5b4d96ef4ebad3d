"""Generate tests/golden/proesmans_golden.npz from the REFERENCE: pysteps/motion/proesmans.py
driving the reference's own _proesmans.pyx, compiled out of tree with the reference flags
(-fopenmp -O3 -ffast-math, setup.py:27-28) exactly like _vet.pyx (see gen_vet_golden.py):

    B=/tmp/vetbuild; cp /root/reference/pysteps/motion/_proesmans.pyx $B/pysteps/motion/
    (setup: Extension("pysteps.motion._proesmans", ..., extra_compile_args=["-fopenmp","-O3","-ffast-math"]))
    cd $B && CC=/usr/bin/gcc LDSHARED="/usr/bin/gcc -shared" python setup_p.py build_ext --inplace
    python tests/golden/gen_proesmans_golden.py /tmp/vetbuild
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from proesmans_cases import CASES, build_case  # noqa: E402


def main():
    build = sys.argv[1] if len(sys.argv) > 1 else "/tmp/vetbuild"
    ref = ref_module("pysteps.motion.proesmans", build)
    out = {}
    for name in CASES:
        frames, kw = build_case(name)
        if frames.shape[0] != 2:
            try:
                ref.proesmans(frames, **kw)
            except ValueError as e:
                out[name + "/error"] = np.array(str(e))
            frames = frames[-2:]
        adv, q = ref.proesmans(frames, full_output=True, **kw)
        out[name + "/advfield"] = adv
        out[name + "/quality"] = q
        assert np.array_equal(ref.proesmans(frames, **kw), adv[0])
    path = os.path.join(HERE, "proesmans_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
