#!/bin/bash
set -u
mkdir -p gpurun_out
for N in 4 2; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-parity > gpurun_out/r2_scale_N$N.json 2> gpurun_out/r2_scale_N$N.err
echo "N=$N rc=$?"; tail -c 150 gpurun_out/r2_scale_N$N.json; tail -2 gpurun_out/r2_scale_N$N.err | cut -c1-200
done
