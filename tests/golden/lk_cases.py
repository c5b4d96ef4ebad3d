"""Seeded Lucas-Kanade test cases shared by the golden generator and the tests."""
import numpy as np

from pysteps_b200 import _synthetic as syn

CASES = ["plain_160x200", "nan_200x176", "three_frames_192x160", "odd_width_150x203"]


def build_case(name):
    """-> (input_images (T,m,n) float64 [NaN = no data], kwargs)."""
    if name == "plain_160x200":
        return syn.rain_frames(160, 200, 2, 0), {}
    if name == "nan_200x176":
        fr = syn.rain_frames(200, 176, 2, 1)
        fr = np.stack([syn.nan_disc(f, 0.2) for f in fr])
        fr[:, :12, :] = np.nan
        fr[:, :, -9:] = np.nan
        return fr, {}
    if name == "three_frames_192x160":
        return syn.rain_frames(192, 160, 3, 2, dx=2, dy=3), {}
    if name == "odd_width_150x203":
        # dBR-like field (negative no-rain value), width not a multiple of 32
        fr = syn.rain_frames(150, 203, 2, 3, dx=-2, dy=1)
        fr = np.where(fr > 0.1, 10.0 * np.log10(np.maximum(fr, 0.1)), -15.0)
        return fr, {}
    raise KeyError(name)
