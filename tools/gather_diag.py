"""Diagnostic (torchrun): cost of the pieces of the tile-partitioned composite step."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pysteps_b200 import _shard
import pysteps_b200

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bench.set_workload("composite4096")
M, N = bench.M, bench.N_
frames_h, precip_h, V_h = bench.make_inputs(rank, True)
frames_d = torch.from_numpy(frames_h).cuda(); precip_d = torch.from_numpy(precip_h).cuda()
motion = pysteps_b200.motion.get_method("lk"); extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
band = _shard.row_band(M, world, rank)


def timed(name, fn, reps=5):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    ev = []
    t0 = time.perf_counter()
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(); e.record(); ev.append((s, e))
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / reps * 1e3
    dev = sum(s.elapsed_time(e) for s, e in ev) / reps
    if rank == 0:
        print(f"{name:34s} device {dev:8.3f} ms   host {host:8.3f} ms", flush=True)
    return r


Vband = timed("motion band", lambda: motion(frames_d, interp_kwargs={"b200_rows": band}))
Vfull = timed("motion full", lambda: motion(frames_d))
timed("gather (all_gather list)", lambda: _shard.gather_row_bands(Vband, M, world, rank))
full = torch.empty((2, M, N), dtype=torch.float64, device="cuda")


def g2():
    for c in range(2):
        dist.all_gather_into_tensor(full[c], Vband[c])
    return full


timed("gather (into_tensor, prealloc)", g2)
timed("broadcast full", lambda: dist.broadcast(Vfull, src=0))
timed("extrap band", lambda: extrap(precip_d, Vfull, bench.T_LEAD, b200_rows=band))
timed("step", lambda: extrap(precip_d, _shard.gather_row_bands(motion(frames_d, interp_kwargs={"b200_rows": band}), M, world, rank), bench.T_LEAD, b200_rows=band))
dist.destroy_process_group()
