"""Summarise an ncu launch list as a markdown table (and the Schur kernel's DRAM bytes per launch as JSON).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 \
        --csv --log-file X.csv python scripts/profile_step.py c2          (on the GPU box, NGT_B200_NO_GRAPH=1)
    python scripts/launch_list_summary.py X.csv [steps_in_file=2] [traffic.json]

The last step's launches are tabulated (earlier ones are warm-up)."""
import csv, json, sys
from collections import OrderedDict

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
traffic_out = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, mi, ui, idi = (h.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "ID"))
launches = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    d = launches.setdefault(r[idi], {"name": r[ki].split("(")[0].replace("void ", "").replace("ngt::", "")})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3)
    else:
        d[r[mi]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
L = list(launches.values())
step = L[len(L) - len(L) // nsteps:]
agg = OrderedDict()
for d in step:
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["us"]
    a[2] = max(a[2], d["us"])
    a[3] += d.get("dram__bytes_read.sum", 0)
    a[4] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print("launches in the step: %d, sum of durations %.2f ms\n" % (len(step), tot / 1e3))
print("| kernel | launches | sum us | avg us | max us | share | DRAM read MB | DRAM write MB |")
print("|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f | %.1f%% | %.1f | %.1f |" % (name, a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot,
                                                                     a[3] / 1e6, a[4] / 1e6))
if traffic_out:
    up = [d for d in step if d["name"].startswith("update_dmma")]
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in up)
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in up)
    json.dump({"c2_bytes_per_launch": (rd + wr) / len(up), "launches": len(up), "dram_read_bytes": rd, "dram_write_bytes": wr,
               "kernel_us_under_ncu": sum(d["us"] for d in up),
               "source": "update_dmma_kernel<128> and <64> rows of the ncu launch list profiles/r01_launches_c2.csv "
                         "(second of two steps of scripts/profile_step.py c2, NGT_B200_NO_GRAPH=1), "
                         "scripts/launch_list_summary.py"}, open(traffic_out, "w"), indent=1)
