"""Time the opt-in paths on a B200 (run by tools/round2_first_call.sh with the PYSTEPS_B200_*
switches set): spline orders of the extrapolator, Proesmans, exact-ties LK."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _lib
from pysteps_b200 import _synthetic as syn


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    with _lib.Trace() as tr:
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
    stages = {k: round(sum(v) / reps, 3) for k, v in tr.summary().items()}
    print(json.dumps({"what": name, "ms": round(ms, 3), "stage_ms": stages}), flush=True)


m = n = 2048
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
V = torch.from_numpy(syn.velocity_field(m, n, 0)).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
timed("extrapolate order 1, T=12, 2048^2", lambda: extrap(P, V, 12))
if os.environ.get("PYSTEPS_B200_ENABLE_SPLINE") == "1":
    for order, mode in ((3, "constant"), (3, "nearest"), (0, "constant"), (5, "nearest")):
        timed(f"extrapolate order {order} {mode}, T=12, 2048^2",
              lambda: extrap(P, V, 12, interp_order=order, map_coordinates_mode=mode))
frames = torch.from_numpy(syn.rain_frames(m, n, 3, 0, dx=3, dy=-2)).cuda()
lk = pysteps_b200.motion.get_method("lk")
timed("dense_lucaskanade 3 frames 2048^2 (exact ties: %s)" % os.environ.get("PYSTEPS_B200_EXACT_TIES", "0"),
      lambda: lk(frames))
if os.environ.get("PYSTEPS_B200_ENABLE_PROESMANS") == "1":
    pro = pysteps_b200.motion.get_method("proesmans")
    for size, it in ((512, 100), (1024, 100), (2048, 20)):
        fr = torch.from_numpy(syn.rain_frames(size, size, 2, 0, dx=3, dy=-2)).cuda()
        timed(f"proesmans {size}^2, 6 levels, {it} iterations", lambda: pro(fr, num_iter=it), reps=1)
