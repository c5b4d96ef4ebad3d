"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST
of `reps` repetitions."""
import collections
import csv
import re
import sys

path, reps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = [(x["Kernel Name"], float(x["Metric Value"])) for x in csv.DictReader(lines)
        if x.get("Metric Name") == "gpu__time_duration.sum"]
n = len(recs) // reps
agg = collections.OrderedDict()
for k, v in recs[-n:]:
    a = agg.setdefault(re.sub(r"\(.*", "", k), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print(f"{k[:56]:56s} {c:3d} {t / 1000:8.1f} us {100 * t / tot:5.1f}%")
print(n, "launches", round(tot / 1000, 1), "us")
