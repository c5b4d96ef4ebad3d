"""CUDA-event time of the semi-Lagrangian extrapolation call (T = 12, 2048^2) for three kinds of
advection field.  B200_SL_INTOPS=0 selects the kernel variant without the integer-pipe tricks."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _lib
from pysteps_b200 import _synthetic as syn

m = n = 2048
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
lk = pysteps_b200.motion.get_method("lk")
fields = {"smooth": torch.from_numpy(syn.velocity_field(m, n, 0, "smooth")).cuda(),
          "rotation": torch.from_numpy(syn.velocity_field(m, n, 0, "rotation") * 2.0).cuda(),
          "lk": lk(torch.from_numpy(syn.rain_frames(m, n, 2, 0)).cuda())}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
out = {}
for name, V in fields.items():
    for _ in range(5):
        extrap(P, V, 12)
    ms = []
    for _ in range(20):
        flush.fill_(1)
        with _lib.Trace(only=("b200_sl_extrapolate_rows",)) as tr:
            extrap(P, V, 12)
        ms += tr.summary()["b200_sl_extrapolate_rows"]
    ms.sort()
    out[name] = {"median_ms": round(ms[len(ms) // 2], 4), "min_ms": round(ms[0], 4)}
print(json.dumps({"variant": "intops=" + os.environ.get("B200_SL_INTOPS", "1"), "kernel_ms": out}))
