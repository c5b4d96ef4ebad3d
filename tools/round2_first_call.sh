#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root, 1 GPU):
# the default suite, then every opt-in path (built and CPU-verified this round, never run on
# hardware) with its tests and timings.  Results -> gpurun_out/round2_*.log
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/round2_default_tests.log
PYSTEPS_B200_ENABLE_SPLINE=1 timeout 400 python -m pytest tests/test_spline_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/round2_spline_tests.log
PYSTEPS_B200_ENABLE_PROESMANS=1 timeout 400 python -m pytest tests/test_proesmans_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/round2_proesmans_tests.log
PYSTEPS_B200_EXACT_TIES=1 timeout 400 python -m pytest tests/test_lk_gpu.py -m gpu -q -k exact_ties 2>&1 | tail -15 | tee gpurun_out/round2_exact_ties_tests.log
PYSTEPS_B200_ENABLE_SPLINE=1 PYSTEPS_B200_ENABLE_PROESMANS=1 timeout 420 python tools/optin_timing.py 2>&1 | tail -20 | tee gpurun_out/round2_optin_timing.log
PYSTEPS_B200_EXACT_TIES=1 timeout 300 python tools/optin_timing.py 2>&1 | grep lucaskanade | tee -a gpurun_out/round2_optin_timing.log
