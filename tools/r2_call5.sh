#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_sl_gpu.py tests/test_vet_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c5_tests.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -8 | tee gpurun_out/r2c5_lk_timing.log
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c5_lk_launches.csv python tools/lk_once.py > gpurun_out/r2c5_lk_once.log 2>&1
timeout 600 python tools/vet_time.py 2>&1 | tail -4 | tee gpurun_out/r2c5_vet_time.log
timeout 900 python bench.py --workload vet_sl12_2048 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r2c5_bench_vet.json 2> gpurun_out/r2c5_bench_vet.err; tail -c 400 gpurun_out/r2c5_bench_vet.json; tail -3 gpurun_out/r2c5_bench_vet.err
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -k "vet" 2>&1 | tail -8 | tee gpurun_out/r2c5_vet_baseline.log
