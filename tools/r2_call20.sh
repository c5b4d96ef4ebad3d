#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2c20_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err; tail -c 200 gpurun_out/r2c20_bench.json; tail -3 gpurun_out/r2c20_bench.err
