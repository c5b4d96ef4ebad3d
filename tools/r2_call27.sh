#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_reference_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2c27_tests.log
timeout 600 python tools/lk_timing.py > gpurun_out/r2c27_lk_timing.log 2>&1; grep "2 frames" gpurun_out/r2c27_lk_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c27_launches_lk.csv python tools/lk_once.py > gpurun_out/r2c27_ncu.log 2>&1; tail -1 gpurun_out/r2c27_ncu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; tail -c 300 gpurun_out/r2c27_bench.json; tail -3 gpurun_out/r2c27_bench.err
