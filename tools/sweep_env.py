#!/usr/bin/env python
"""Run bench.py once per value of an environment variable and print one line each:
    python tools/sweep_env.py B200_SL_BY 8 4 16 2
(tuning aid; bench.py --no-cpu, 20 steps)."""
import json
import os
import subprocess
import sys

var, values = sys.argv[1], sys.argv[2:]
for v in values:
    env = dict(os.environ)
    env[var] = v
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "5", "--no-cpu"],
                         capture_output=True, text=True, env=env).stdout.strip().splitlines()
    d = json.loads(out[-1])
    st = d["stage_ms_per_step"]
    print(var, v, "ms/step", round(d["ms_per_step"], 3), "sl_kernel_ms", round(d["roofline"]["kernel_ms"], 4),
          "idw", round(st.get("b200_idw_fill", 0), 3), "decluster", round(st.get("b200_decluster", 0), 3),
          flush=True)
