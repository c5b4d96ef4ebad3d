"""Seeded BPS (velocity perturbation) cases shared by the golden generator and the tests."""
import numpy as np

from pysteps_b200 import _synthetic as syn

M, N = 56, 72
LEADS = (5.0, 10.0, 37.5)        # minutes
MEMBERS = ((11, 1.0), (12, 2.0))   # (seed, kmperpixel)
TIMESTEP = 5.0


def fields(kind):
    P = syn.rain_field(M, N, 2)
    V = 3.0 * syn.velocity_field(M, N, 2, "rotation" if kind == "rotation" else "smooth")
    if kind == "zeros":        # calm patches: |V| <= 1e-12 -> unperturbed there
        V = V.copy()
        V[:, 10:20, 30:50] = 0.0
        V[:, 40, 5] = 1e-13
    if kind == "float32":
        V = V.astype(np.float32)
    return P, V


KINDS = ("smooth", "rotation", "zeros", "float32")
