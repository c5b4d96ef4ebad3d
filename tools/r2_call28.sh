#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export B200_NO_CLUSTER_SORT=1; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-parity > gpurun_out/r2c28_bench_$v.json 2> gpurun_out/r2c28_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2c28_bench_$v.json').read().strip().splitlines()[-1])
print('nocluster=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_step']['b200_good_features'])
PY
done
