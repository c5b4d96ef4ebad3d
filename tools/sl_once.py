"""extrapolate(P, V, 12) at 2048^2 a few times (for ncu captures).  FIELD=smooth|rotation|lk"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _synthetic as syn

m = n = 2048
kind = os.environ.get("FIELD", "lk")
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
if kind == "lk":
    V = pysteps_b200.motion.get_method("lk")(torch.from_numpy(syn.rain_frames(m, n, 2, 0)).cuda())
else:
    V = torch.from_numpy(syn.velocity_field(m, n, 0, kind) * (2.0 if kind == "rotation" else 1.0)).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
for _ in range(int(os.environ.get("REPS", "3"))):
    out = extrap(P, V, 12)
torch.cuda.synchronize()
