"""Where does a member-step's time go?  Device-resident member loop at 2048^2 for M members."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pysteps_b200
from pysteps_b200 import _lib
from pysteps_b200 import _synthetic as syn

m = n = 2048
T = 12
extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
bps_init, bps_gen = pysteps_b200.noise.get_method("bps")
V = torch.from_numpy(syn.velocity_field(m, n, 0)).cuda()
P = torch.from_numpy(syn.rain_field(m, n, 0).astype(np.float32)).cuda()


def loop(M):
    fields = [P * (1.0 + 0.01 * i) for i in range(M)]
    perts = [bps_init(V, 1.0, 5.0, randstate=np.random.RandomState(1000 + i)) for i in range(M)]
    disp = [None] * M
    for t in range(T):
        for j in range(M):
            Vm = V + bps_gen(perts[j], (t + 1) * 5.0)
            _, disp[j] = extrap(fields[j], Vm, [1.0], displacement_prev=disp[j], return_displacement=True,
                                b200_resident=False)


for M in (6, 12, 24, 24):
    loop(M)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(M)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"M={M}: {1e3 * dt:.1f} ms, {1e6 * dt / (M * T):.0f} us per member-step, "
          f"allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB", flush=True)
with _lib.Trace() as tr:
    loop(24)
print({k: (len(v), round(sum(v), 2)) for k, v in tr.summary().items()})
pr = cProfile.Profile()
pr.enable()
loop(24)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:4500])
