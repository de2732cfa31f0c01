"""profiles/r02_traffic.json from the ncu launch list of one BASELINE config 2 step (this build):
    python scripts/make_traffic_json.py <launches_c2.csv> <steps_in_file> <pool_bytes> "<command>"
schur_bytes_per_launch = DRAM bytes (read + write) of the update_dmma kernels / their launches; step_dram_bytes = DRAM
bytes of every kernel of the step + the pool memset (not a kernel, so not in the list).  bench.py quotes these numbers
only while the sources' fingerprint still matches."""
import csv, json, os, sys
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

path, nsteps, pool_bytes, command = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, mi, ui, idi = (h.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "ID"))
launches = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    d = launches.setdefault(r[idi], {"name": r[ki]})
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        d["us"] = v / 1e3
    else:
        d[r[mi]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
L = list(launches.values())
step = L[len(L) - len(L) // nsteps:]
byt = lambda d: d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
up = [d for d in step if "update_dmma" in d["name"]]
out = {
    "build_fingerprint": bench.build_fingerprint(), "command": command,
    "schur_bytes_per_launch": sum(byt(d) for d in up) / max(1, len(up)), "schur_launches": len(up),
    "schur_dram_bytes": sum(byt(d) for d in up), "schur_us_under_ncu": sum(d["us"] for d in up),
    "kernel_dram_bytes": sum(byt(d) for d in step), "pool_memset_bytes": pool_bytes,
    "step_dram_bytes": sum(byt(d) for d in step) + pool_bytes, "launches_in_step": len(step),
    "sum_kernel_us_under_ncu": sum(d["us"] for d in step),
    "note": "per-launch numbers are cold-cache (ncu flushes the caches before every kernel) and serialised: an upper bound "
            "for the traffic of the CUDA-graph replay that bench.py times",
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
