"""Diagnostic: where the end-to-end (NumPy API) time of the default workload goes."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pysteps_b200

bench.set_workload("lk_sl12_2048")
frames_h, precip_h, V_h = bench.make_inputs(0, True)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
frames_h, precip_h = pin(frames_h), pin(precip_h)
motion = pysteps_b200.motion.get_method("lk"); extrap = pysteps_b200.extrapolation.get_method("semilagrangian")


def timed(name, fn, reps=10, nbytes=None):
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    extra = f"  {nbytes / ms / 1e6:7.1f} GB/s" if nbytes else ""
    print(f"{name:40s} {ms:8.3f} ms{extra}", flush=True)
    return r


ft = torch.from_numpy(frames_h)
d = timed("H2D frames (pinned, 100 MB)", lambda: ft.to("cuda", non_blocking=True), nbytes=frames_h.nbytes)
big = torch.empty((12, 2048, 2048), dtype=torch.float32, device="cuda")
host = torch.empty(big.shape, dtype=big.dtype, pin_memory=True)
timed("D2H 12 fields (pinned, reused, 201 MB)", lambda: host.copy_(big, non_blocking=True), nbytes=big.numel() * 4)
from pysteps_b200 import _device
timed("D2H 12 fields (_device.to_host)", lambda: _device.to_host(big), nbytes=big.numel() * 4)
Vh = timed("motion(frames_h) -> NumPy", lambda: motion(frames_h))
Vd = timed("motion(frames_d) -> device", lambda: motion(d))
timed("extrap(precip_h, V_h, 12) -> NumPy", lambda: extrap(precip_h, Vh, 12))
pd = torch.from_numpy(precip_h).cuda()
timed("extrap(precip_d, V_d, 12) -> device", lambda: extrap(pd, Vd, 12))
timed("whole step NumPy", lambda: extrap(precip_h, motion(frames_h), 12))
