"""H2D / D2H bandwidth of the paths the NumPy API uses (pinned buffers through torch)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysteps_b200 import _device

_device.require_cuda()
for mb in (16, 64, 192):
    n = mb * 1024 * 1024 // 8
    hp = torch.empty(n, dtype=torch.float64, pin_memory=True)
    hnp = hp.numpy()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    pageable = np.ones(n)

    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return mb / 1024 / ((time.perf_counter() - t0) / reps)

    print(f"{mb} MB: H2D pinned tensor copy_ {t(lambda: d.copy_(hp, non_blocking=True)):.1f} GB/s | "
          f"H2D to_device(numpy view of pinned) {t(lambda: _device.to_device(hnp)):.1f} | "
          f"H2D to_device(pageable numpy) {t(lambda: _device.to_device(pageable)):.1f} | "
          f"D2H copy_ into pinned {t(lambda: hp.copy_(d, non_blocking=True)):.1f} | "
          f"D2H to_host (alloc pinned + copy + sync) {t(lambda: _device.to_host(d)):.1f}", flush=True)
