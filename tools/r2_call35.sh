#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity']['sparse_bit_identical'], d['parity']['extrapolation_bit_identical'])
PY
