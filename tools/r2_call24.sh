#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sl_gpu.py -m gpu -x -q -s 2>&1 | tail -6 | tee gpurun_out/r2c24_tests.log
timeout 600 python tools/sl_f32_timing.py > gpurun_out/r2c24_sl_f32.json 2> gpurun_out/r2c24_sl_f32.err; cat gpurun_out/r2c24_sl_f32.json; tail -3 gpurun_out/r2c24_sl_f32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c24_launches_slf32.csv python tools/sl_f32_once.py > gpurun_out/r2c24_ncu.log 2>&1; tail -2 gpurun_out/r2c24_ncu.log
