#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lk_gpu.py -m gpu -x -q -k "certificate or api_behaviour or golden" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['ms_each_step_rank0'])
e=d['ensemble24']; print('ens', e['ms_per_step'], e.get('member_loop_ms'), e.get('batched_ms_per_step'), e['ms_each_step_rank0'])
c=d['composite4096']; print('comp', c['ms_per_step'], c['ms_each_step_rank0'])
PY
