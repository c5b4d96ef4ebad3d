"""The opt-in float32-tap trajectory kernel beside the exact one (T = 12, 2048^2, three kinds of advection
field): CUDA-event time of each C call (the float32 call includes its float64 -> float32 copy of the
velocity), share of pixels recomputed by the exact fallback, and the largest deviations."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _lib
from pysteps_b200 import _synthetic as syn

m = n = int(os.environ.get("SIZE", "2048"))
T = 12
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
lk = pysteps_b200.motion.get_method("lk")
fields = {"smooth": torch.from_numpy(syn.velocity_field(m, n, 0, "smooth")).cuda(),
          "rotation": torch.from_numpy(syn.velocity_field(m, n, 0, "rotation") * 2.0).cuda(),
          "lk": lk(torch.from_numpy(syn.rain_frames(m, n, 2, 0)).cuda())}
# the benchmark's synthetic motion is (3, -2) px per step: displacements cluster on INTEGERS, i.e. on cell
# boundaries, which is the worst case for certifying the floor.  The same field with a fractional mean:
off = torch.tensor([0.37, 0.21], dtype=torch.float64, device="cuda").view(2, 1, 1)
fields["smooth_fractional_mean"] = fields["smooth"] + off
fields["lk_fractional_mean"] = fields["lk"] + off
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(call, name, **kw):
    for _ in range(5):
        extrap(P, V, T, **kw)
    ms = []
    for _ in range(20):
        flush.fill_(1)
        with _lib.Trace(only=(call,)) as tr:
            extrap(P, V, T, **kw)
        ms += tr.summary()[call]
    ms.sort()
    return {"median_ms": round(ms[len(ms) // 2], 4), "min_ms": round(ms[0], 4)}


out = {}
for name, V in fields.items():
    exact, dex = extrap(P, V, T, return_displacement=True)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    fast, dfa = extrap(P, V, T, return_displacement=True, b200_float32_taps=True, b200_fallback_count=cnt)
    err = torch.nan_to_num((fast.double() - exact.double()).abs(), nan=0.0).max().item()
    out[name] = {
        "exact": timed("b200_sl_extrapolate_rows", name),
        "float32_taps": timed("b200_sl_extrapolate_rows_f32", name, b200_float32_taps=True),
        "recomputed_fraction": cnt.item() / (m * n),
        "max_value_error_over_max_precip": err / float(P.abs().max()),
        "max_displacement_error_px": (dfa - dex).abs().max().item(),
        "nan_pattern_equal": bool(torch.equal(torch.isnan(fast), torch.isnan(exact))),
    }
print(json.dumps({"size": [m, n], "T": T, "fields": out}))
