#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sl_gpu.py tests/test_bps_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_lk_gpu.py -m gpu -x -q -k "fused or stages or float32_frames" 2>&1 | tail -2
timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee gpurun_out/r02_sl_timing.json
timeout 300 python tools/sl_f32_timing.py 2>/dev/null | tail -1 | tee gpurun_out/r02_sl_f32_timing.json | cut -c1-400
