#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ens_diag2.py > gpurun_out/r2c30_ens.txt 2>&1; head -70 gpurun_out/r2c30_ens.txt | cut -c1-170
