#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_gputests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r02_smoke.log
