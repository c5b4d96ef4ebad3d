"""Per-stage CUDA-event times of dense_lucaskanade on 3 frames of 2048^2 (device-resident)."""
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


def timed(name, fn, reps=10):
    for _ in range(3):
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


m = n = int(os.environ.get("SIZE", "2048"))
frames = torch.from_numpy(syn.rain_frames(m, n, 3, 0, dx=3, dy=-2)).cuda()
lk = pysteps_b200.motion.get_method("lk")
timed(f"dense_lucaskanade 3 frames {m}^2", lambda: lk(frames))
frames2 = torch.from_numpy(syn.rain_frames(m, n, 2, 0, dx=3, dy=-2)).cuda()
timed(f"dense_lucaskanade 2 frames {m}^2", lambda: lk(frames2))
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
for kind in ("smooth", "rotation"):
    V = torch.from_numpy(syn.velocity_field(m, n, 0, kind) * (2.0 if kind == "rotation" else 1.0)).cuda()
    timed(f"extrapolate T=12 {kind} f64 field {m}^2", lambda: extrap(P, V, 12))
Vlk = lk(frames)
timed(f"extrapolate T=12 LK field {m}^2", lambda: extrap(P, Vlk, 12))
