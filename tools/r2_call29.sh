#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/host_profile.py > gpurun_out/r2c29_host_profile.txt 2>&1; head -60 gpurun_out/r2c29_host_profile.txt
