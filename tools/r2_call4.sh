#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_sl_gpu.py tests/test_bps_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c4_tests.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -8 | tee gpurun_out/r2c4_lk_timing.log
timeout 120 python tools/pcie_diag.py 2>&1 | tail -4 | tee gpurun_out/r2c4_pcie.log
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c4_lk_launches.csv python tools/lk_once.py > gpurun_out/r2c4_lk_once.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; tail -c 600 gpurun_out/r2c4_bench.json; tail -5 gpurun_out/r2c4_bench.err
timeout 1500 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2c4_baseline_tests.log
