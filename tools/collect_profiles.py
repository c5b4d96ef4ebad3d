#!/usr/bin/env python
"""Turn the artefacts of tools/final_run.sh (in gpurun_out/) into the tracked summaries
under profiles/:   python tools/collect_profiles.py r01"""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
d = json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])
json.dump(d, open(f"profiles/{tag}_bench_line.json", "w"), indent=1)
r = json.loads(open(f"gpurun_out/bench_{tag}_ref.json").read().strip().splitlines()[-1])
json.dump(r, open(f"profiles/{tag}_bench_reference_arm.json", "w"), indent=1)
subprocess.check_call([sys.executable, "tools/ncu_summary.py", f"gpurun_out/{tag}_kernels.ncu-rep",
                       f"profiles/{tag}_final_kernels"], stdout=subprocess.DEVNULL)
subprocess.check_call([sys.executable, "tools/ncu_summary.py", f"gpurun_out/{tag}_kernels.ncu-rep",
                       "/tmp/slonly", "sl_multistep"], stdout=subprocess.DEVNULL)
json.dump(json.load(open("/tmp/slonly_traffic.json")), open("profiles/sl_traffic.json", "w"), indent=1)
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = [i for i, r_ in enumerate(rows) if r_ and r_[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 2:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r_ in data:
    if len(r_) <= vi:
        continue
    name = r_[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg.setdefault(name, []).append(float(r_[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(f"profiles/{tag}_launches_final.md", "w") as f:
    f.write(f"# ncu launch list of the {tag} final step (LK + 12-leadtime semi-Lagrangian, 2048^2)\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 python bench.py "
            "--steps 2 --warmup 3 --no-cpu`\n(140 launches ~ 2 steps; cold-cache serialised launches: "
            "compare shares, not absolutes)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| {k[:70]} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n")
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "ref", r["value"])
