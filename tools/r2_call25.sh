#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sl_gpu.py -m gpu -x -q -k float32 2>&1 | tail -3 | tee gpurun_out/r2c25_tests.log
timeout 300 python tools/sl_f32_timing.py > gpurun_out/r2c25_sl_f32.json 2> gpurun_out/r2c25_sl_f32.err; cat gpurun_out/r2c25_sl_f32.json; tail -2 gpurun_out/r2c25_sl_f32.err
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lk_track_kernel|kd_build_kernel|front_kernel" -s 7 -c 7 -f -o gpurun_out/r2c25_lk python tools/lk_once.py > gpurun_out/r2c25_ncu.log 2>&1; tail -2 gpurun_out/r2c25_ncu.log
