#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-parity > gpurun_out/r2_scale_N${N}b.json 2> gpurun_out/r2_scale_N${N}b.err
echo "rc=$?"; tail -c 200 gpurun_out/r2_scale_N${N}b.json; tail -3 gpurun_out/r2_scale_N${N}b.err | cut -c1-200
