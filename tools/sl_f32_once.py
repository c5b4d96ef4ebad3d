"""One exact and one float32-tap extrapolation call (2048^2, T = 12, LK-like smooth field) for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _synthetic as syn

m = n = 2048
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
V = torch.from_numpy(syn.velocity_field(m, n, 0, os.environ.get("FIELD", "smooth"))).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
for _ in range(int(os.environ.get("REPS", "2"))):
    extrap(P, V, 12)
    extrap(P, V, 12, b200_float32_taps=True)
torch.cuda.synchronize()
