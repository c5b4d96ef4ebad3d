#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_knn_gpu.py tests/test_sl_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c10_tests.log
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_reference_gpu.py -m gpu -x -q -k "lk or lucaskanade or composite or steps" 2>&1 | tail -6 | tee gpurun_out/r2c10_baseline.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -8 | tee gpurun_out/r2c10_lk_timing.log
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c10_lk_launches.csv python tools/lk_once.py > gpurun_out/r2c10_lk_once.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; tail -c 300 gpurun_out/r2c10_bench.json; tail -3 gpurun_out/r2c10_bench.err
