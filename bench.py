#!/usr/bin/env python
"""bench.py -- Mpix/s advected on BASELINE.json's config[1]:
2048x2048 synthetic radar frames, motion field + 12-leadtime semi-Lagrangian
extrapolation.

    python bench.py --gpus N --steps K --warmup W          (ours, CUDA)
    python bench.py --impl reference ...                   (CPU oracle port, host cores)

One "step" = one pass of the hot path over one batch of synthetic input, on every rank:
  motion field from this rank's last three frames (dense Lucas-Kanade),
  12-leadtime semi-Lagrangian extrapolation of this rank's field with it.
Metric: Mpix/s advected = (N * T * m * n) / step time, max over ranks (weak scaling: N
independent nowcasts, one per GPU, no collective on the data path).
Other workloads (--workload): ensemble24 (24 BPS-perturbed members round-robin over the GPUs,
one NCCL broadcast of the motion field), composite4096 (one 4096^2 composite, output row bands
over the GPUs, NCCL all-gather of the motion-field bands), vet_sl12_2048.

value  : inputs resident in HBM, device time by CUDA events, L2 flushed between steps.
e2e    : the same step through the public NumPy API with pinned HOST buffers, H2D and
         D2H copies inside the timed region.
roofline: the semi-Lagrangian trajectory kernel (sl_multistep_kernel), algorithmic
         bytes per launch / CUDA-event time of that launch, vs MEASURED_PEAKS.json.
cpu_baseline: the CPU oracle (port of the reference path) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M = N_ = 2048
T_LEAD = 12
METRIC = "Mpix/s advected (2048^2 frame, 12 leadtimes)"
UNIT = "Mpix/s"
MOTION = "lk"
SCALING = "weak"
MEMBERS = 0

# BASELINE.json configs.  The default (config[1]) is what `metric` is quoted on; the others are
# optional extra measurements (python bench.py --workload ...).
WORKLOADS = {
    "lk_sl12_2048": dict(m=2048, n=2048, T=12, motion="lk", scaling="weak",
                         metric="Mpix/s advected (2048^2 frame, 12 leadtimes)"),
    "vet_sl12_2048": dict(m=2048, n=2048, T=12, motion="vet", scaling="weak",
                          metric="Mpix/s advected (2048^2 frame, 12 leadtimes, VET motion)"),
    "composite4096": dict(m=4096, n=4096, T=24, motion="lk", scaling="strong",
                          metric="Mpix/s advected (4096^2 composite, 24 leadtimes, row bands over GPUs)"),
    # config[3]'s advection component: the call shape of nowcasts/utils.py:441-458 -- every
    # member has its own (perturbed) velocity, precipitation field and carried displacement,
    # one single-step extrapolator call per member and lead time; members sharded over ranks
    "ensemble24": dict(m=2048, n=2048, T=12, motion="lk", scaling="strong", members=24,
                       metric="Mpix/s advected (24-member ensemble, 2048^2, 12 single-step calls per member)"),
}


def set_workload(name):
    global M, N_, T_LEAD, METRIC, MOTION, SCALING
    w = WORKLOADS[name]
    M, N_, T_LEAD, METRIC, MOTION, SCALING = w["m"], w["n"], w["T"], w["metric"], w["motion"], w["scaling"]
    global MEMBERS
    MEMBERS = w.get("members", 0)


def have_lk():
    try:
        from pysteps_b200.motion import lucaskanade  # noqa: F401
        return True
    except ImportError:
        return False


def workload_name(lk):
    mot = {"lk": "lk_dense", "vet": "vet"}[MOTION] if lk else "given_field"
    if MEMBERS:
        return f"{mot}+semilagrangian_{MEMBERS}members_x{T_LEAD}single_steps_{M}x{N_}"
    return f"{mot}+semilagrangian_T{T_LEAD}_{M}x{N_}" + ("_rowbands" if SCALING == "strong" else "")


def make_inputs(seed, lk):
    """frames: float64 (2,m,n) for the motion estimator (as pysteps importers deliver them);
    precip: the last frame as float32 (the synthetic data are float32-exact; a float32 array
    keeps the reference's float64 arithmetic and halves the output volume);
    V: synthetic float32 advection field, used only when the LK stage is not built."""
    from pysteps_b200 import _synthetic as syn
    frames = syn.rain_frames(M, N_, 2, seed)
    V = syn.velocity_field(M, N_, seed).astype(np.float32)
    return frames, frames[-1].astype(np.float32), V


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi in loop mode (-lms 200, the profiling recipe's clocks line) from the warm-up to the
    end of the end-to-end leg.  The first sample (taken before any load) is dropped."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, enabled=True):
        self.index = index
        # one sampler per job (rank 0's GPU): nvidia-smi polls serialise on a driver-wide lock
        # that kernel launches of every process on the box also take
        self.enabled = enabled and not os.environ.get("BENCH_NO_CLOCKS")
        self.samples = []
        self.proc = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)  # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in (out or "").splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)
        if len(self.samples) >= 3:
            self.samples = self.samples[1:]

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names)
                   if any(len(s) > 2 + k and s[2 + k].lower() == "active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def stage_rooflines(trace_ms, steps, m, n, peak_gbs, sl_bytes):
    """Per C-ABI call of the step: algorithmic bytes per call (SURVEY.md section 8d figures, with
    the float64 frames this implementation keeps), achieved GB/s over the CUDA-event time of
    the call, fraction of the HBM peak, and what actually bounds it (from the ncu captures in
    profiles/).  `trace_ms` = {name: [ms of every call in the timed region]}."""
    px = m * n
    table = {
        "b200_idw_fill": (16 * px, "alu/latency: exhaustive-in-tile k-NN with register-resident sorted lists"),
        "b200_sl_extrapolate_rows": (sl_bytes, "L1 wavefronts + FP64 pipe (reference-order float64 trajectories)"),
        "b200_min_eig": (5 * px, "latency: one sequential FP64 running sum per column (OpenCV-exact box filter)"),
        "b200_quantise_u8": (10 * px, "hbm/l2 streaming"),
        "b200_mask_invalid": (9 * px, "hbm streaming + reduction"),
        "b200_morph_opening": (17 * px, "hbm/l2 streaming stencil"),
        "b200_masked_minmax": (9 * px, "hbm streaming reduction"),
        "b200_lk_track": (None, "latency: ordered float32 window sums, one CTA per feature"),
        "b200_good_features": (None, "latency: sort + ordered greedy selection"),
        "b200_bps_perturb_velocity": (32 * px, "hbm streaming"),
    }
    out = []
    for name, ms in trace_ms.items():
        if name not in table or not ms:
            continue
        nbytes, bound = table[name]
        avg = sum(ms) / len(ms)
        row = {"call": name, "ms_per_call": avg, "calls_per_step": len(ms) / steps,
               "algorithmic_bytes_per_call": nbytes, "bound": bound}
        if nbytes is not None and avg > 0:
            row["achieved_gbs"] = nbytes / (avg * 1e-3) / 1e9
            row["frac_of_hbm_peak"] = row["achieved_gbs"] / peak_gbs
        out.append(row)
    out.sort(key=lambda r: -r["ms_per_call"] * r["calls_per_step"])
    return out


# ----------------------------------------------------------------------------- CPU legs
def cpu_step(frames, precip, V, lk):
    """The oracle port of one step on host cores."""
    from oracle import semilagrangian as ora
    if lk and MOTION == "vet":
        from oracle import vet as ora_vet
        V = ora_vet.vet(frames, verbose=False)
    elif lk:
        from oracle import lucaskanade as ora_lk
        V = ora_lk.dense_lucaskanade(frames)
    if MEMBERS:
        # bounded sample: ONE member of the ensemble (the members are independent and identical
        # in cost); the caller extrapolates to MEMBERS members
        from oracle import noise_motion as ora_bps
        t1 = time.perf_counter()
        pert = ora_bps.initialize_bps(V, 1.0, 5.0, randstate=np.random.RandomState(1000))
        disp, res = None, None
        for t in range(T_LEAD):
            Vm = ora_bps.perturbed_velocity(V, pert, (t + 1) * 5.0)
            res, disp = ora.extrapolate(precip, Vm, [1.0], displacement_prev=disp, return_displacement=True)
        return res, time.perf_counter() - t1
    return ora.extrapolate(precip, V, T_LEAD)


def _timed_cpu_step(frames, precip, V, lk):
    """seconds of one whole step on the host (ensemble: motion + MEMBERS x the sampled member)"""
    t0 = time.perf_counter()
    r = cpu_step(frames, precip, V, lk)
    dt = time.perf_counter() - t0
    if MEMBERS:
        member = r[1]
        return (dt - member) + MEMBERS * member
    return dt


def cpu_baseline(frames, precip, V, lk, reps=1):
    import oracle
    best = None
    for _ in range(reps):
        dt = _timed_cpu_step(frames, precip, V, lk)
        best = dt if best is None else min(best, dt)
    return {"value": (MEMBERS or 1) * T_LEAD * M * N_ / best / 1e6, "unit": UNIT, "cores": oracle.num_threads(),
            "kind": "port",
            "sample": f"{reps} full step(s) of {workload_name(lk)} (oracle C/NumPy port, "
                      f"OpenMP {oracle.num_threads()} threads), best of {reps}; {best:.2f} s"
                      + (f" (1 of {MEMBERS} members run, scaled)" if MEMBERS else "")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lk = have_lk_oracle()
    frames, precip, V = make_inputs(0, lk)
    import oracle
    for _ in range(args.warmup):
        cpu_step(np.ascontiguousarray(frames[:, :256, :256]), np.ascontiguousarray(precip[:256, :256]),
                 np.ascontiguousarray(V[:, :256, :256]), lk)
    # Every step is a bounded sample of the workload: the full frame when the host is fast enough
    # for the whole run to end within a few minutes (64 cores: 2.5 s per step), else a centred
    # crop sized from the first step's time; the metric is per advected pixel either way.
    budget_s = float(os.environ.get("BENCH_REFERENCE_BUDGET_S", "240"))
    dt = _timed_cpu_step(frames, precip, V, lk)
    pixels = M * N_
    sample = f"full {M}x{N_} frame"
    side_m, side_n = M, N_
    if dt * args.steps > budget_s and args.steps > 1:
        frac = max((budget_s - dt) / (dt * (args.steps - 1)), 1.0 / 64.0)
        side_m = max(256, int(M * frac ** 0.5) // 32 * 32)
        side_n = max(256, int(N_ * frac ** 0.5) // 32 * 32)
        r0, c0 = (M - side_m) // 2, (N_ - side_n) // 2
        frames = np.ascontiguousarray(frames[:, r0:r0 + side_m, c0:c0 + side_n])
        precip = np.ascontiguousarray(precip[r0:r0 + side_m, c0:c0 + side_n])
        V = np.ascontiguousarray(V[:, r0:r0 + side_m, c0:c0 + side_n])
        sample = f"first step on the full frame, the others on a centred {side_m}x{side_n} crop"
    for _ in range(args.steps - 1):
        dt += _timed_cpu_step(frames, precip, V, lk)
        pixels += side_m * side_n
    val = (MEMBERS or 1) * T_LEAD * pixels / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(lk), "frame": [M, N_], "leadtimes": T_LEAD},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": oracle.num_threads(),
                             "kind": "port",
                             "sample": f"{args.steps} step(s), {sample}; oracle port of the reference "
                                       f"path, OpenMP {oracle.num_threads()} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def have_lk_oracle():
    try:
        from oracle import lucaskanade  # noqa: F401
        return have_lk()
    except ImportError:
        return False


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import pysteps_b200
    from pysteps_b200 import _lib, _shard
    extrap = pysteps_b200.extrapolation.get_method("semilagrangian")
    lk = have_lk()
    motion = None
    if lk:
        from pysteps_b200 import motion as b200_motion
        motion = b200_motion.get_method(MOTION)
        if MOTION == "vet":
            _vet = motion
            motion = lambda fr: _vet(fr.cpu().numpy() if torch.is_tensor(fr) else fr, verbose=False)  # noqa: E731

    frames_h, precip_h, V_h = make_inputs(rank, lk)
    # pinned host buffers for the e2e leg
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    frames_h = pin(frames_h)
    precip_h = pin(precip_h)
    V_h = pin(V_h)
    frames_d = torch.from_numpy(frames_h).cuda()
    precip_d = torch.from_numpy(precip_h).cuda()
    V_d = torch.from_numpy(V_h).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # strong scaling (one composite): every rank owns a band of output rows, inputs replicated
    band = _shard.row_band(M, world, rank) if SCALING == "strong" else None
    ekw = {} if band is None else {"b200_rows": band}

    def step_device():
        """inputs resident in HBM; results stay in HBM."""
        if lk and band is not None and world > 1 and MOTION == "lk":
            # one composite over GPUs: the sparse LK stages are deterministic and replicated, every
            # rank fills its band of the motion field, the bands are all-gathered (the exchange
            # step of this path), and every rank advects its band of output rows
            Vband = motion(frames_d, interp_kwargs={"b200_rows": band})
            Vd = _shard.gather_row_bands(Vband, M, world, rank)
            return extrap(precip_d, Vd, T_LEAD, **ekw)
        if lk and SCALING == "weak":
            # independent nowcasts: every rank estimates the motion of ITS frames and advects ITS
            # field -- no collective on the data path
            Vd = motion(frames_d)
            if not torch.is_tensor(Vd):
                Vd = torch.from_numpy(np.ascontiguousarray(Vd)).cuda()
            return extrap(precip_d, Vd, T_LEAD)
        if lk:
            if rank == 0:
                Vd = motion(frames_d)
                if not torch.is_tensor(Vd):
                    Vd = torch.from_numpy(np.ascontiguousarray(Vd)).cuda()
            else:
                Vd = torch.empty((2, M, N_), dtype=torch.float64, device="cuda")
        else:
            Vd = V_d
        Vd = _shard.broadcast_field(Vd, src=0)  # the only collective (NCCL over NVLink)
        return extrap(precip_d, Vd, T_LEAD, **ekw)

    if MEMBERS:
        from pysteps_b200 import noise as b200_noise
        bps_init, bps_gen = b200_noise.get_method("bps")
        mine = _shard.member_indices(MEMBERS, world, rank)
        # per-member precipitation fields (synthetic); the velocity perturbations are the BPS
        # perturbator's (noise/motion.py), one seeded RandomState per member as in
        # nowcasts/steps.py:915-926, evaluated inside the advection call (fused kernel)
        member_precip = [precip_d * (1.0 + 0.01 * i) for i in mine]
        member_precip_h = [pin(precip_h * np.float32(1.0 + 0.01 * i)) for i in mine]
        KMPP, DT_MIN = 1.0, 5.0

        def member_loop(V, fields, resident):
            """nowcasts/utils.py:440-458: T lead times x this rank's members, each a single-step
            call with a freshly perturbed motion field, carrying its own displacement."""
            perts = [bps_init(V, 1.0 / KMPP, DT_MIN, randstate=np.random.RandomState(1000 + i)) for i in mine]
            disp = [None] * len(mine)
            last = None
            for t in range(T_LEAD):
                for j in range(len(mine)):
                    Vm = V + bps_gen(perts[j], (t + 1) * DT_MIN)
                    last, disp[j] = extrap(fields[j], Vm, [1.0], displacement_prev=disp[j],
                                           return_displacement=True, b200_resident=resident)
            return last

        def step_device():  # noqa: F811
            """rank 0: motion field; broadcast; member loop on device-resident fields."""
            if rank == 0:
                Vd = motion(frames_d)
            else:
                Vd = torch.empty((2, M, N_), dtype=torch.float64, device="cuda")
            Vd = _shard.broadcast_field(Vd, src=0)
            return member_loop(Vd, member_precip, False)

    def step_host():
        """public NumPy API: H2D of inputs and D2H of the result inside."""
        if MEMBERS:
            # NumPy API as nowcast_main_loop uses it: the motion field is uploaded once per
            # forecast, every member-step uploads its precipitation field and downloads the
            # advected one; perturbed fields and displacements never leave the device
            Vh = motion(frames_h) if rank == 0 else None
            if world > 1:
                Vd = torch.from_numpy(Vh).cuda() if rank == 0 else \
                    torch.empty((2, M, N_), dtype=torch.float64, device="cuda")
                dist.broadcast(Vd, src=0)
                Vh = Vd.cpu().numpy()
            return member_loop(Vh, member_precip_h, True)
        if lk and band is not None and world > 1 and MOTION == "lk":
            Vband = motion(frames_h, interp_kwargs={"b200_rows": band})  # NumPy band
            Vh = _shard.gather_row_bands(torch.from_numpy(Vband).cuda(), M, world, rank).cpu().numpy()
        elif lk and SCALING == "weak":
            Vh = motion(frames_h)  # NumPy (2,m,n) float64, as pysteps returns; independent per rank
        elif lk:
            Vh = motion(frames_h) if rank == 0 else None
            if world > 1:
                Vd = torch.from_numpy(Vh).cuda() if rank == 0 else \
                    torch.empty((2, M, N_), dtype=torch.float64, device="cuda")
                dist.broadcast(Vd, src=0)
                Vh = Vd.cpu().numpy()
        else:
            Vh = V_h
        return extrap(precip_h, Vh, T_LEAD, **ekw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------
    # The clock sampler (nvidia-smi -lms 200, one per job) spans the warm-up and the timed steps and
    # is stopped before the end-to-end leg: every poll holds a driver lock that launches and copies
    # also take (a 20 ms period was measured to slow a 6 ms step by 3 ms, a 200 ms period the
    # 10.7 ms end-to-end step by 3.7 ms).  The warm-up is stretched to >= 0.5 s of the same load
    # so that the median is over several samples under load although the timed region is short.
    clocks = ClockSampler(local, enabled=(rank == 0))
    clocks.__enter__()
    t_w, n_w = time.perf_counter(), 0
    while True:
        step_device()
        torch.cuda.synchronize()
        n_w += 1
        done = n_w >= args.warmup and (n_w >= 400 or time.perf_counter() - t_w >= 0.5)
        if world > 1:  # rank 0 decides, so every rank runs the same number of steps
            flag = torch.tensor([int(done)], device="cuda")
            dist.broadcast(flag, src=0)
            done = bool(flag.item())
        if done:
            break
    barrier()
    launches0 = _lib.load().b200_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    with _lib.Trace() as trace:
        for s, e in ev:
            flush.fill_(1)  # L2 flush between timed iterations (outside the events)
            barrier()
            s.record()
            step_device()
            e.record()
        barrier()
    clocks.__exit__(None, None, None)
    launches = _lib.load().b200_launch_count() - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    tr = trace.summary()
    dev_ms = _shard.max_over_ranks(dev_ms, device="cuda")
    nfields = world if SCALING == "weak" else 1
    if MEMBERS:
        nfields = MEMBERS
    value = nfields * args.steps * T_LEAD * M * N_ / (dev_ms * 1e-3) / 1e6

    # ---- end-to-end timing (host buffers) --------------------------------------------
    # the pinned host pool (torch's caching host allocator: one cudaHostAlloc per new block, tens of
    # ms for a 200 MB result) reaches its steady state after a few calls
    for _ in range(max(6, args.warmup)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("BENCH_E2E_BREAKDOWN") and lk and not MEMBERS:  # diagnostic only
        tm = te = 0.0
        for _ in range(args.steps):
            a = time.perf_counter(); Vx = motion(frames_h); b = time.perf_counter()
            out = extrap(precip_h, Vx, T_LEAD); c = time.perf_counter()
            tm += b - a; te += c - b
        print(f"e2e breakdown: loop {1e3 * e2e_s / args.steps:.2f} ms/step; motion {1e3 * tm / args.steps:.2f} "
              f"extrap {1e3 * te / args.steps:.2f}", file=sys.stderr)
    e2e_s = _shard.max_over_ranks(e2e_s, device="cuda")
    e2e_val = nfields * args.steps * T_LEAD * M * N_ / e2e_s / 1e6
    # LK: frames up, field down, field up again for the extrapolator (plugin API is NumPy)
    h2d = precip_h.nbytes + (frames_h.nbytes + 2 * M * N_ * 8 if lk else V_h.nbytes)
    d2h = out.nbytes + (2 * M * N_ * 8 if lk else 0)
    if MEMBERS:  # per rank: frames + field once, one precipitation field up / one down per member-step
        h2d = frames_h.nbytes + 2 * M * N_ * 8 + len(mine) * T_LEAD * precip_h.nbytes
        d2h = 2 * M * N_ * 8 + len(mine) * T_LEAD * precip_h.nbytes

    if rank == 0:
        # ---- roofline of the trajectory kernel (this run's launches) -----------------
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
        k_ms = tr.get("b200_sl_extrapolate_rows", []) or tr.get("b200_sl_extrapolate", [])
        k_avg = sum(k_ms) / len(k_ms) if k_ms else float("nan")
        vbytes = 8 if lk else 4  # LK returns float64 fields, synthetic V is float32
        rows_here = M if band is None else band[1] - band[0]
        alg_bytes = M * N_ * (2 * vbytes + 4) + rows_here * N_ * 4 * T_LEAD
        if MEMBERS:  # single-step member calls: V 16 + precip 4 + displacement in/out 32 + out 4 B per pixel
            alg_bytes = M * N_ * 56
        achieved = alg_bytes / (k_avg * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "sl_traffic.json")))["bytes_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": "sl_multistep_kernel", "bound": "hbm", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "kernel_ms": k_avg,
                    "algorithmic_bytes_per_launch": alg_bytes}
        stage_ms = {k: sum(v) / args.steps for k, v in tr.items()}
        try:
            stages = stage_rooflines(tr, args.steps, M, N_, peak, alg_bytes)
        except Exception as exc:  # supplementary table only
            stages = [{"error": repr(exc)}]
        # the CPU leg is a property of the host, measured once: rank 0 of the single-GPU run only
        cpu = cpu_baseline(frames_h, precip_h, V_h, have_lk_oracle()) if (not args.no_cpu and world == 1) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
                "dtype": "f64 (trajectories, motion field, IDW); f32 precip in/out; u8/i16/f32 "
                         "OpenCV-exact LK stages", "data": "synthetic",
                "config": {"workload": workload_name(lk), "frame": [M, N_], "leadtimes": T_LEAD,
                           "fields_per_gpu": 1 if SCALING == "weak" else round(1.0 / world, 4), "l2": "flushed between timed steps (256 MB fill)",
                           "parallelism": (
                               f"{world} independent nowcast(s), one per GPU, no collective" if SCALING == "weak" else
                               f"{MEMBERS} members round-robin over {world} GPU(s), NCCL broadcast of the motion field"
                               if MEMBERS else
                               f"output row bands over {world} GPU(s), inputs replicated"
                               + (", NCCL all-gather of the motion-field bands" if world > 1 and MOTION == "lk"
                                  else ""))},
                "clocks": clocks.summary(),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "stage_ms_per_step": stage_ms,
                "stage_rooflines": stages}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="lk_sl12_2048", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
