#!/bin/bash
mkdir -p gpurun_out
timeout 120 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-parity > gpurun_out/r2c36_bench.json 2> gpurun_out/r2c36_bench.err; tail -2 gpurun_out/r2c36_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c36_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], d['ms_each_step_rank0'])
PY
