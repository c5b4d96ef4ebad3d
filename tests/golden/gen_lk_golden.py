"""Generate tests/golden/lk_golden.npz by running the REFERENCE
(pysteps/motion/lucaskanade.py and the helpers it calls, with the cv2 4.13.0 and
scipy 1.18.1 binaries of this container) stage by stage.

    python tests/golden/gen_lk_golden.py

Stored per case: the corners of the first frame pair (feature/shitomasi.py), the tracked
vectors (tracking/lucaskanade.py), the pooled outlier mask and declustered vectors
(utils/cleansing.py) and the dense field (utils/interpolate.py); for one case also the
raw cv2 intermediates (minimum-eigenvalue map, pyramid, Scharr derivatives).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refimport import ref_module  # noqa: E402
from lk_cases import CASES, build_case  # noqa: E402


def main():
    import cv2
    lk = ref_module("pysteps.motion.lucaskanade")
    sh = ref_module("pysteps.feature.shitomasi")
    cl = ref_module("pysteps.utils.cleansing")
    im = ref_module("pysteps.utils.images")
    tr = ref_module("pysteps.tracking.lucaskanade")
    out = {}
    for name in CASES:
        frames, kw = build_case(name)
        a = np.ma.masked_invalid(frames[0]); np.ma.set_fill_value(a, a.min())
        b = np.ma.masked_invalid(frames[1]); np.ma.set_fill_value(b, b.min())
        a = im.morph_opening(a, a.min(), 3)
        b = im.morph_opening(b, b.min(), 3)
        out[name + "/opened0"] = a.filled(-9999.0).astype(np.float32)
        pts = sh.detection(a).astype(np.float32)
        out[name + "/points"] = pts
        xy, uv = tr.track_features(a, b, pts)
        out[name + "/xy"] = xy
        out[name + "/uv"] = uv
        sxy, suv = lk.dense_lucaskanade(frames, dense=False, **kw)
        out[name + "/sparse_xy"] = sxy
        out[name + "/sparse_uv"] = suv
        dxy, duv = cl.decluster(sxy, suv, 20, 1)
        out[name + "/decl_xy"] = dxy
        out[name + "/decl_uv"] = duv
        out[name + "/dense"] = lk.dense_lucaskanade(frames, **kw)
    # raw OpenCV intermediates for one quantised frame
    frames, _ = build_case("odd_width_150x203")
    q = ((frames[0] - frames[0].min()) / (frames[0].max() - frames[0].min()) * 255).astype(np.uint8)
    out["cv/q"] = q
    out["cv/min_eig"] = cv2.cornerMinEigenVal(q, 5, ksize=3)
    nl, pyr = cv2.buildOpticalFlowPyramid(q, (21, 21), 3, withDerivatives=True)
    for l in range(nl + 1):
        out[f"cv/pyr{l}"] = np.ascontiguousarray(pyr[2 * l])
        out[f"cv/deriv{l}"] = np.ascontiguousarray(pyr[2 * l + 1])
    path = os.path.join(HERE, "lk_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
