"""bench.py's ensemble24 device step at N = 1 under cProfile: where the non-kernel time goes."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

b = bench.Bench()
w = dict(bench.WORKLOADS["ensemble24"], name="ensemble24")
step_device, step_host, info = bench.build_ensemble(b, w)
for _ in range(3):
    step_device()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    print(f"step_device: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
with b.lib.Trace() as tr:
    t0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    print(f"step_device under Trace: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
print({k: (len(v), round(sum(v), 2)) for k, v in tr.summary().items()})
b.flush.fill_(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
step_device()
torch.cuda.synchronize()
print(f"step_device after an L2 flush: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
step_device()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4000])
