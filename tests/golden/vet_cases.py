"""Seeded VET test cases shared by the golden generator and the tests."""
import numpy as np

from pysteps_b200 import _synthetic as syn

EVAL_CASES = ["eval_128x160_s4x4", "eval_128x160_s16x8", "eval_256x256_s32x32", "eval_96x96_s2x2",
              "eval_3frames_120x144_s8x4"]
FIELD_CASES = ["field_128x128", "field_3frames_nan_200x168", "field_padding_127x150"]


def eval_case(name):
    """-> (sector_displacement (2,xs,ys), images (T,nx,ny), mask int8 (nx,ny), smooth_gain)."""
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    spec = {"eval_128x160_s4x4": (128, 160, 4, 4, 2), "eval_128x160_s16x8": (128, 160, 16, 8, 2),
            "eval_256x256_s32x32": (256, 256, 32, 32, 2), "eval_96x96_s2x2": (96, 96, 2, 2, 2),
            "eval_3frames_120x144_s8x4": (120, 144, 8, 4, 3)}[name]
    m, n, xs, ys, T = spec
    fr = syn.rain_frames(m, n, T, 1)
    mask = np.zeros((m, n), np.int8)
    mask[10:30, 40:70] = 1
    mask[:, :3] = 1
    sd = np.ascontiguousarray(rng.normal(size=(2, xs, ys)) * 2.5)
    return sd, np.ascontiguousarray(fr), mask, 1e6


def field_case(name):
    """-> (input_images, kwargs) for vet()."""
    if name == "field_128x128":
        return syn.rain_frames(128, 128, 2, 2), {}
    if name == "field_3frames_nan_200x168":
        fr = syn.rain_frames(200, 168, 3, 2)
        fr[:, 50:70, 60:90] = np.nan
        return fr, {}
    if name == "field_padding_127x150":
        return syn.rain_frames(127, 150, 2, 2), {"padding": 3}
    raise KeyError(name)
