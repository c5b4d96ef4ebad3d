#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sl_gpu.py tests/test_bps_gpu.py tests/test_spline_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2c12_tests.log
B200_SL_INTOPS=0 timeout 900 python -m pytest tests/test_sl_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee -a gpurun_out/r2c12_tests.log
timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee gpurun_out/r2c12_sl_timing.log
B200_SL_INTOPS=0 timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee -a gpurun_out/r2c12_sl_timing.log
timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee -a gpurun_out/r2c12_sl_timing.log
B200_SL_INTOPS=0 timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee -a gpurun_out/r2c12_sl_timing.log
