#!/bin/bash
# bench.py as the driver launches it at N GPUs (ours + reference arm); results -> gpurun_out/r2_scale_N*.json
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_N$N.json 2> gpurun_out/r2_scale_N$N.err
echo "rc=$?"; tail -c 300 gpurun_out/r2_scale_N$N.json; tail -5 gpurun_out/r2_scale_N$N.err
if [ "${2:-}" = "ref" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_scale_ref_N$N.json 2> gpurun_out/r2_scale_ref_N$N.err
echo "rc=$?"; tail -c 600 gpurun_out/r2_scale_ref_N$N.json
fi
