"""Seeded Proesmans test cases shared by the golden generator and the tests."""
import numpy as np

from pysteps_b200 import _synthetic as syn

CASES = ["default_96x128", "levels3_lam10_70x55", "dbr_3frames_64x64", "tiny_pyramid_40x40", "constant_32x32"]


def build_case(name):
    """-> (input_images, kwargs)"""
    if name == "default_96x128":
        return syn.rain_frames(96, 128, 2, 0, dx=2, dy=-1), {"num_iter": 40}
    if name == "levels3_lam10_70x55":
        return syn.rain_frames(70, 55, 2, 1, dx=-1, dy=2), {"num_levels": 3, "lam": 10.0, "num_iter": 25}
    if name == "dbr_3frames_64x64":
        fr = syn.rain_frames(64, 64, 3, 2, dx=1, dy=1)
        return np.where(fr > 0.1, 10.0 * np.log10(np.maximum(fr, 0.1)), -15.0), {"num_levels": 4, "num_iter": 30}
    if name == "tiny_pyramid_40x40":
        return syn.rain_frames(40, 40, 2, 3, dx=1, dy=0), {"num_levels": 6, "num_iter": 10}
    if name == "constant_32x32":
        return np.full((2, 32, 32), 3.5), {"num_levels": 3, "num_iter": 5}
    if name == "illconditioned_63x95":
        # one level, strong data term, many iterations: rounding-level perturbations grow to pixels
        return syn.rain_frames(63, 95, 2, 7, dx=3, dy=-2), {"num_levels": 1, "lam": 1000.0, "num_iter": 100}
    if name == "odd_sizes_5levels_173x77":
        return syn.rain_frames(173, 77, 2, 8, dx=2, dy=0), {"num_levels": 5, "num_iter": 100}
    raise KeyError(name)

# compared bit for bit with the reference source built without -ffast-math
STRICT_CASES = CASES + ["illconditioned_63x95", "odd_sizes_5levels_173x77"]
