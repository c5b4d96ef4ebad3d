"""CPU tests of bench.py's host-side arithmetic (no GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_stage_rooflines_table():
    m = n = 2048
    trace = {"b200_idw_fill": [0.9, 0.8], "b200_sl_extrapolate_rows": [0.5, 0.4],
             "b200_lk_track": [0.4, 0.4, 0.4, 0.4], "b200_compact_rows": [0.01, 0.01], "b200_min_eig": []}
    rows = bench.stage_rooflines(trace, 2, m, n, 6576.7, 285212672)
    assert [r["call"] for r in rows] == ["b200_idw_fill", "b200_lk_track", "b200_sl_extrapolate_rows"]
    idw = rows[0]
    assert idw["calls_per_step"] == 1 and abs(idw["ms_per_call"] - 0.85) < 1e-12
    assert idw["algorithmic_bytes_per_call"] == 16 * m * n
    assert abs(idw["achieved_gbs"] - 16 * m * n / 0.85e-3 / 1e9) < 1e-9
    assert abs(idw["frac_of_hbm_peak"] - idw["achieved_gbs"] / 6576.7) < 1e-15
    assert rows[1]["calls_per_step"] == 2 and "achieved_gbs" not in rows[1]
    json.dumps(rows)


def test_workloads_and_metric_names():
    for name, w in bench.WORKLOADS.items():
        bench.set_workload(name)
        assert bench.M == w["m"] and bench.T_LEAD == w["T"] and bench.SCALING in ("weak", "strong")
        assert "Mpix/s" in bench.METRIC
    w = bench.set_workload("lk_sl12_2048")
    assert bench.MEMBERS == 0 and bench.workload_name(True) == "lk_dense+semilagrangian_T12_2048x2048"
    # both arms print the same config dict (the driver compares them)
    assert bench.config_of(w) == {"workload": "lk_dense+semilagrangian_T12_2048x2048", "frame": [2048, 2048],
                                  "leadtimes": 12}
    assert bench.workload_name(True, dict(bench.WORKLOADS["ensemble24"])) == \
        "lk_dense+semilagrangian_24members_x12single_steps_2048x2048"


def test_reference_arm_runs_with_all_host_threads_under_a_launcher(monkeypatch, capsys):
    """torchrun exports OMP_NUM_THREADS=1; the reference arm sets the thread count explicitly and
    reports it, on the same config as the CUDA arm."""
    import types
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    monkeypatch.setenv("BENCH_REFERENCE_BUDGET_S", "1")
    w = dict(bench.WORKLOADS["lk_sl12_2048"], name="tiny", m=256, n=256, T=2)
    args = types.SimpleNamespace(steps=2, warmup=0, gpus=4)
    assert bench.run_reference(args, w) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.config_of(w)
    ncpu = os.cpu_count() or 1
    assert line["cpu_baseline"]["cores"] in (ncpu, max(ncpu // 2, 1)) and line["cpu_baseline"]["cores"] > 1
    assert line["n_gpus"] == 4 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]


def test_clock_sampler_is_silent_inside_a_timed_region():
    """bench.py's NVML sampler: the background thread polls only while not paused; inside a timed region the
    main thread samples at step boundaries (a background query there stalled one timed step by 50 ms)."""
    import time

    import bench

    class FakeNvml:
        NVML_CLOCK_SM = 0
        calls = 0

        def nvmlDeviceGetClockInfo(self, h, k):
            FakeNvml.calls += 1
            return 1965

        def nvmlDeviceGetMaxClockInfo(self, h, k):
            return 1965

        def nvmlDeviceGetCurrentClocksEventReasons(self, h):
            return 0x4  # sw_power_cap: kept and noted

    c = bench.ClockSampler(0, enabled=True)
    c._nvml_handle = lambda: (FakeNvml(), object())
    c.__enter__()
    time.sleep(0.3)
    c.paused = True
    time.sleep(0.25)
    n = FakeNvml.calls
    time.sleep(0.45)
    assert FakeNvml.calls == n, "the background thread touched NVML while paused"
    c.sample_now()
    c.sample_now()
    assert FakeNvml.calls == n + 2
    c.paused = False
    c.__exit__(None, None, None)
    s = c.summary()
    assert s["sm_mhz"] == 1965 and s["reasons"] == ["sw_power_cap"] and s["samples"] >= 3
    off = bench.ClockSampler(0, enabled=False)
    off.sample_now()  # no handle: a no-op on the ranks that do not sample
    assert off.summary()["reasons"] == ["unsampled"]


def test_settle_gc_freezes_the_heap():
    import gc

    import bench
    junk = [[i] for i in range(1000)]
    before = gc.get_freeze_count()
    bench.settle_gc()
    assert gc.get_freeze_count() > before
    del junk
    gc.unfreeze()
