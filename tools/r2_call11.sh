#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sl_gpu.py tests/test_bps_gpu.py tests/test_spline_gpu.py tests/test_lk_gpu.py tests/test_vet_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c11_tests.log
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c11_baseline.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -8 | tee gpurun_out/r2c11_lk_timing.log
timeout 300 python tools/vet_time.py 2>&1 | tail -3 | tee gpurun_out/r2c11_vet_time.log
REPS=2 FIELD=lk timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_multistep -s 1 -c 1 -f -o gpurun_out/r2_sl2_full python tools/sl_once.py > gpurun_out/r2c11_ncu_sl.log 2>&1; tail -2 gpurun_out/r2c11_ncu_sl.log
