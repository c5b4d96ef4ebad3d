#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bps_gpu.py tests/test_sl_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2c15_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; tail -c 300 gpurun_out/r2c15_bench.json; tail -3 gpurun_out/r2c15_bench.err
