"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) as a markdown table.
    python scripts/launch_list_summary.py X.csv [steps_in_file=2]
The last step's launches are tabulated (earlier ones are warm-up)."""
import csv, sys
from collections import OrderedDict

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("nsecond", "ns"):
        v /= 1e3
    elif r[ui] in ("msecond", "ms"):
        v *= 1e3
    name = r[ki].split("(")[0].replace("void ", "").replace("ngt::", "")
    seq.append((name, r[gi], v))
n = len(seq) // nsteps
step = seq[len(seq) - n:]
agg = OrderedDict()
for name, _g, us in step:
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
print("launches in the step: %d, sum of durations %.2f ms" % (len(step), tot / 1e3))
print("| kernel | launches | sum us | avg us | max us | share |")
print("|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f | %.1f%% |" % (name, a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
