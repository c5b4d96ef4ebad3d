#!/bin/bash
# usage (under gpurun --gpus N): bash tools/scale_run.sh N [workload ...]
N=$1
shift
W=${@:-lk_sl12_2048 ensemble24 composite4096}
for w in $W; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --workload $w 2>/dev/null | tail -1 > gpurun_out/scale_${w}_n${N}.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/scale_${w}_n${N}.json").read())
print("N=${N}", d["config"]["workload"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["scaling"], "e2e", round(d["e2e"]["value"], 1))
print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
PY
done
