#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sl_gpu.py tests/test_bps_gpu.py tests/test_spline_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c7_tests.log
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -x -q -k "extrapolation or composite" 2>&1 | tail -6 | tee gpurun_out/r2c7_baseline.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -3 | tee gpurun_out/r2c7_sl_timing.log
timeout 300 python tools/ens_diag.py 2>&1 | tail -60 | tee gpurun_out/r2c7_ens_diag.log
