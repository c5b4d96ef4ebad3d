#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_reference_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2c26_tests.log
timeout 600 python tools/lk_timing.py > gpurun_out/r2c26_lk_timing.log 2>&1; grep "2 frames" gpurun_out/r2c26_lk_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c26_launches_lk.csv python tools/lk_once.py > gpurun_out/r2c26_ncu.log 2>&1; tail -1 gpurun_out/r2c26_ncu.log
timeout 300 python tools/sl_f32_timing.py > gpurun_out/r2c26_sl_f32.json 2> gpurun_out/r2c26_sl_f32.err; cat gpurun_out/r2c26_sl_f32.json; tail -2 gpurun_out/r2c26_sl_f32.err
