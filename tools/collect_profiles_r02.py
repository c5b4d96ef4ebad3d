#!/usr/bin/env python
"""Turn the artefacts of tools/final_run_r02.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

T = "r02"
G = "gpurun_out"


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def launches(csv_path, md_path, title, reps=3):
    lines = [ln for ln in open(csv_path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    n = len(rows) // reps
    last = rows[(reps - 1) * n:]
    agg = collections.OrderedDict()
    for r in last:
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg.setdefault(name, []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(md_path, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over {reps} repetitions; the "
                f"last one is listed ({n} launches, {tot / 1e3:.1f} us in total).  Cold-cache, serialised launches: "
                "compare SHARES, not absolutes.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k[:80]} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n")


def main():
    d = last_json(f"{G}/bench_{T}.json")
    json.dump(d, open(f"profiles/{T}_bench_line.json", "w"), indent=1)
    if os.path.exists(f"{G}/bench_{T}_ref.json"):
        json.dump(last_json(f"{G}/bench_{T}_ref.json"), open(f"profiles/{T}_bench_reference_arm.json", "w"), indent=1)
    if os.path.exists(f"{G}/bench_{T}_vet.json"):
        json.dump(last_json(f"{G}/bench_{T}_vet.json"), open(f"profiles/{T}_bench_vet_line.json", "w"), indent=1)
    for n in (2, 4, 8):
        p = f"{G}/r2_scale_N{n}.json"
        if os.path.exists(p):
            json.dump(last_json(p), open(f"profiles/{T}_scale_N{n}.json", "w"), indent=1)
        p = f"{G}/r2_scale_ref_N{n}.json"
        if os.path.exists(p):
            json.dump(last_json(p), open(f"profiles/{T}_scale_ref_N{n}.json", "w"), indent=1)
    launches(f"{G}/{T}_launches_lk.csv", f"profiles/{T}_launches_lk.md",
             "ncu launch list: dense_lucaskanade, 2 frames of 2048^2 (tools/lk_once.py)")
    launches(f"{G}/{T}_launches_sl.csv", f"profiles/{T}_launches_sl.md",
             "ncu launch list: LK field + extrapolate(P, V, 12) at 2048^2 (tools/sl_once.py)")
    if os.path.exists(f"{G}/{T}_launches_sl_f32.csv"):
        launches(f"{G}/{T}_launches_sl_f32.csv", f"profiles/{T}_launches_sl_f32.md",
                 "ncu launch list: one exact and one float32-tap extrapolate(P, V, 12) at 2048^2 (tools/sl_f32_once.py)",
                 reps=2)
    for rep, out, sub in ((f"{G}/{T}_sl.ncu-rep", f"profiles/{T}_sl_final", "sl_multistep"),
                          (f"{G}/{T}_sl_f32.ncu-rep", f"profiles/{T}_sl_f32", "sl_f32_kernel"),
                          (f"{G}/{T}_lk.ncu-rep", f"profiles/{T}_lk_kernels", ""),
                          (f"{G}/{T}_vet.ncu-rep", f"profiles/{T}_vet_eval", "vet_eval")):
        if os.path.exists(rep):
            subprocess.check_call([sys.executable, "tools/ncu_summary.py", rep, out] + ([sub] if sub else []),
                                  stdout=subprocess.DEVNULL)
    if os.path.exists(f"profiles/{T}_sl_final_traffic.json"):
        json.dump(json.load(open(f"profiles/{T}_sl_final_traffic.json")), open("profiles/sl_traffic.json", "w"), indent=1)
    for name in (f"{T}_sl_f32_timing.json", f"{T}_sl_timing.json", f"{T}_lk_timing.log", f"{T}_vet_time.log", f"{T}_gputests.log", f"{T}_smoke.log"):
        if os.path.exists(f"{G}/{name}"):
            open(f"profiles/{name}", "w").write(open(f"{G}/{name}").read())
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "frac", d["roofline"]["frac"])


if __name__ == "__main__":
    main()
