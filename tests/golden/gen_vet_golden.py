"""Generate tests/golden/vet_golden.npz from the REFERENCE: pysteps/motion/vet.py driving
the reference's own _vet.pyx, compiled out of tree with the reference flags
(-fopenmp -O3 -ffast-math, setup.py:27-28):

    B=/tmp/vetbuild; mkdir -p $B/pysteps/motion; cp /root/reference/pysteps/motion/_vet.pyx $B/pysteps/motion/
    (setup.py: Extension("pysteps.motion._vet", ..., extra_compile_args=["-fopenmp","-O3","-ffast-math"]))
    cd $B && CC=/usr/bin/gcc LDSHARED="/usr/bin/gcc -shared" python setup.py build_ext --inplace
    python tests/golden/gen_vet_golden.py /tmp/vetbuild
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from vet_cases import EVAL_CASES, FIELD_CASES, eval_case, field_case  # noqa: E402


def main():
    build = sys.argv[1] if len(sys.argv) > 1 else "/tmp/vetbuild"
    ref = ref_module("pysteps.motion.vet", build)
    out = {}
    for name in EVAL_CASES:
        sd, images, mask, gain = eval_case(name)
        shape = sd.shape[1:]
        out[name + "/cost"] = np.array(ref.vet_cost_function(sd.ravel(), images, shape, mask, gain))
        out[name + "/grad"] = ref.vet_cost_function_gradient(sd.ravel(), images, shape, mask, gain)
    for name in FIELD_CASES:
        images, kw = field_case(name)
        field, steps = ref.vet(images, verbose=False, intermediate_steps=True, **kw)
        out[name + "/field"] = field
        for k, s in enumerate(steps):
            out[name + f"/step{k}"] = np.ascontiguousarray(s)
    img = eval_case("eval_128x160_s4x4")[1][0]
    rng = np.random.default_rng(5)
    disp = rng.normal(size=(2,) + img.shape) * 6
    w, wm, wg = ref.morph(img, disp, gradient=True)
    out["morph/disp"] = disp
    out["morph/image"] = w
    out["morph/mask"] = wm
    out["morph/grad"] = wg
    path = os.path.join(HERE, "vet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
