#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py tests/test_stages_gpu.py tests/test_sl_gpu.py tests/test_bps_gpu.py tests/test_knn_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2c9_tests.log
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_reference_gpu.py -m gpu -x -q -k "not vet" 2>&1 | tail -6 | tee gpurun_out/r2c9_baseline.log
timeout 300 python tools/lk_timing.py 2>&1 | tail -8 | tee gpurun_out/r2c9_lk_timing.log
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c9_lk_launches.csv python tools/lk_once.py > gpurun_out/r2c9_lk_once.log 2>&1
timeout 600 python bench.py --workload ensemble24 --no-extras --no-cpu --no-parity --steps 5 --warmup 3 > gpurun_out/r2c9_bench_ens.json 2> gpurun_out/r2c9_bench_ens.err; tail -c 300 gpurun_out/r2c9_bench_ens.json
BENCH_NO_CLOCKS=1 timeout 600 python bench.py --no-cpu --no-parity --steps 5 --warmup 3 > gpurun_out/r2c9_bench_noclk.json 2> gpurun_out/r2c9_bench_noclk.err; tail -c 300 gpurun_out/r2c9_bench_noclk.json
