#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md and, for
the roofline's `traffic` key, profiles/<name>_traffic.json.

    python tools/ncu_summary.py gpurun_out/sl_r1c.ncu-rep profiles/sl_r1c [kernel-substring]
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * scale.get(unit, 1)


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu summary of `{rep}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun;",
             "per-launch values (ncu serialises and replays: compare shares, not absolutes).", ""]
    traffic = None
    for d in data:
        if sub and sub not in d[ki]:
            continue
        lines.append(f"## {d[ki]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (d[i], units[i])
                lines.append(f"| {k} | {d[i]} | {units[i]} |")
        if "dram__bytes_read.sum" in vals:
            rd = to_bytes(*vals["dram__bytes_read.sum"])
            wr = to_bytes(*vals["dram__bytes_write.sum"])
            traffic = {"kernel": d[ki], "bytes_per_launch": rd + wr, "dram_read": rd,
                       "dram_write": wr, "source": rep}
            lines.append(f"| dram traffic (read+write) | {rd + wr:.0f} | byte |")
        stalls = [(float(d[i]), h) for i, h in enumerate(hdr)
                  if "average_warp_latency_issue_stalled" in h and h.endswith(".ratio")
                  and d[i] not in ("", "n/a")]
        stalls.sort(reverse=True)
        if stalls:
            lines.append("")
            lines.append("Top stall reasons (warp-cycles per issued instruction): " +
                         ", ".join(f"{h.split('issue_stalled_')[1].split('.')[0]}={v:.2f}"
                                   for v, h in stalls[:6]))
        lines.append("")
    # executed-instruction mix from the source page
    src = ncu_csv(rep, "source")
    if len(src) > 2 and "Source" in src[1]:
        h = src[1]
        si, ei = h.index("Source"), h.index("Instructions Executed")
        cnt = collections.Counter()
        tot = 0
        for r in src[2:]:
            try:
                e = int(r[ei])
            except (ValueError, IndexError):
                continue
            toks = r[si].split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
            cnt[op] += e
            tot += e
        lines.append("### executed warp instructions by opcode (first kernel in the report)")
        lines.append("")
        lines.append("| opcode | warp instructions | share |")
        lines.append("|---|---|---|")
        for k, v in cnt.most_common(14):
            lines.append(f"| {k} | {v} | {100.0 * v / tot:.1f}% |")
        lines.append(f"| total | {tot} | |")
    open(outbase + ".md", "w").write("\n".join(lines) + "\n")
    if traffic:
        json.dump(traffic, open(outbase + "_traffic.json", "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
