#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_gpu.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2c8_reference_tests.log
timeout 300 python tools/ens_diag2.py 2>&1 | tail -50 | tee gpurun_out/r2c8_ens_diag2.log
REPS=2 FIELD=lk timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_multistep -s 1 -c 1 -f -o gpurun_out/r2_sl_full python tools/sl_once.py > gpurun_out/r2c8_ncu_sl.log 2>&1; tail -2 gpurun_out/r2c8_ncu_sl.log
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:idw_kernel -s 1 -c 1 -f -o gpurun_out/r2_idw_full python tools/lk_once.py > gpurun_out/r2c8_ncu_idw.log 2>&1; tail -2 gpurun_out/r2c8_ncu_idw.log
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_build -s 2 -c 1 -f -o gpurun_out/r2_kdbuild_full python tools/lk_once.py > gpurun_out/r2c8_ncu_kd.log 2>&1; tail -2 gpurun_out/r2c8_ncu_kd.log
