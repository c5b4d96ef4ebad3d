#!/bin/bash
# Round-end measurement on the GPU box (run under gpurun from the repo root):
#   tests, smoke, default bench (+cpu baseline), reference arm, launch list, ncu full capture.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 200 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"sl_multistep|idw_kernel|lk_track_kernel|box_chain|select_smem|cov_rowsum|outliers_kernel|decluster" \
    -s 10 -c 8 -o gpurun_out/${TAG}_kernels python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1),
      "cpu", round(d["cpu_baseline"]["value"], 2), "frac", round(d["roofline"]["frac"], 4), d["clocks"])
print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
ls -la gpurun_out | tail -6
