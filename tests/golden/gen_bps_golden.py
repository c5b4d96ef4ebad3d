"""Generate tests/golden/bps_golden.npz by running the REFERENCE
(pysteps/noise/motion.py + pysteps/extrapolation/semilagrangian.py) in this container:

    python tests/golden/gen_bps_golden.py

Stored per velocity kind and member: eps, the perturbation fields at LEADS, and the result of
the member loop of pysteps/nowcasts/utils.py:440-458 (three single-step extrapolations of a
perturbed field carrying the displacement)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from bps_cases import KINDS, LEADS, MEMBERS, TIMESTEP, fields  # noqa: E402


def main():
    nm = ref_module("pysteps.noise.motion")
    sl = ref_module("pysteps.extrapolation.semilagrangian")
    out = {}
    for kind in KINDS:
        P, V = fields(kind)
        for seed, kmpp in MEMBERS:
            rs = np.random.RandomState(seed)
            pert = nm.initialize_bps(V, 1.0 / kmpp, TIMESTEP, randstate=rs)
            key = f"{kind}/{seed}"
            out[key + "/eps"] = np.array([pert["eps_par"], pert["eps_perp"], pert["vsf"]])
            out[key + "/V_par"] = pert["V_par"]
            for t in LEADS:
                out[key + f"/pert_{t}"] = nm.generate_bps(pert, t)
            disp, cur = None, P
            for step in range(1, 4):
                Vp = V + nm.generate_bps(pert, step * TIMESTEP)
                res, disp = sl.extrapolate(P, Vp, [1.0], displacement_prev=disp, return_displacement=True)
                cur = res[0]
            out[key + "/advected"] = cur
            out[key + "/disp"] = disp
    path = os.path.join(HERE, "bps_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
