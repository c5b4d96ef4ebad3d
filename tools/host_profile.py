"""Where the HOST spends its time in one device-resident step (dense LK on 2 frames + 12 leadtimes):
cProfile over STEPS steps; the GPU work is asynchronous, so cumulative times are host costs plus the
waits at the step's two read-backs."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _synthetic as syn

m = n = 2048
steps = int(os.environ.get("STEPS", "100"))
frames = torch.from_numpy(syn.rain_frames(m, n, 2, 0, dx=3, dy=-2)).cuda()
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()
lk = pysteps_b200.motion.get_method("lk")
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")


def step():
    V = lk(frames)
    return extrap(P, V, 12)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
print(f"wall per step: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
