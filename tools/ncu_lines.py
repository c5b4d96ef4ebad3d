"""Hot source lines of one kernel in an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index] [top]
Sums the warp-stall samples and executed warp instructions of the SASS under every CUDA source line."""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + pat, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
path, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path, hdr = r[1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[2] != "-":  # per-line aggregate rows carry "-" as their SASS address
        continue
    d = dict(zip(hdr, r))
    try:
        out.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), path.split("/")[-1],
                    int(r[0]), r[1].strip()[:110]))
    except ValueError:
        pass
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"{ts} samples, {ti} warp instructions")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100 * o[0] / ts:5.1f}% smp {100 * o[1] / ti:5.1f}% ins  {o[2]}:{o[3]:<4d} {o[4]}")
