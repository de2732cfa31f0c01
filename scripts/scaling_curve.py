"""SURVEY §8(d) scaling curve of the sparse workload: GPU numeric step, GPU end to end (host arrays -> rates) and the
reference's C++ NGT (oracle/_ref, one core, compute_rates() only) over landscape sizes.
    python scripts/scaling_curve.py [max_reference_side=100]        (GPU box; prints one JSON object)
The reference is only run up to `max_reference_side` (its cost grows ~n^2.1)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle

max_ref_side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rows = []
for side in (50, 100, 200, 316, 632):
    n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B)
    h.compute_rates()
    for _ in range(3):
        h.timed_steps(1)
    reps = 20 if side <= 316 else 5
    step_ms = h.timed_steps(reps) / reps
    st = h.stats()
    kab = float(h.rates()[0])
    h.close()
    t0 = time.perf_counter()
    ne2e = 3
    for _ in range(ne2e):
        g = Handle.from_coo(n, src, dst, k, A, B)
        g.compute_rates()
        _ = g.rates()
        g.close()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / ne2e
    row = {"side": side, "nodes": int(n), "rates": int(src.size), "eliminated": int(st["n_eliminated"]),
           "gpu_step_ms": step_ms, "gpu_step_nodes_per_s": st["n_eliminated"] / (step_ms * 1e-3),
           "gpu_e2e_ms": e2e_ms, "gpu_e2e_nodes_per_s": st["n_eliminated"] / (e2e_ms * 1e-3),
           "fronts": int(st["n_fronts"]), "levels": int(st["n_levels"]), "max_front": int(st["max_front"]),
           "alg_flops": st["alg_flops"], "k_AB": kab}
    if side <= max_ref_side:
        from oracle.cpu_ngt import CpuNGT, best_kind
        kind = best_kind()
        t0 = time.perf_counter()
        c = CpuNGT(n, src, dst, k, A, B, kind=kind)
        t1 = time.perf_counter()
        c.compute_rates()
        t2 = time.perf_counter()
        ref = c.rates()
        row.update({"reference_kind": kind, "reference_construct_s": t1 - t0, "reference_compute_s": t2 - t1,
                    "reference_nodes_per_s": st["n_eliminated"] / (t2 - t1), "rel_err_k_AB": abs(kab - ref[0]) / abs(ref[0])})
    rows.append(row)
    print("side %4d n %7d: step %8.3f ms  e2e %8.1f ms%s" % (side, n, step_ms, e2e_ms,
          ("  reference %.2f s (rel err %.1e)" % (row["reference_compute_s"], row["rel_err_k_AB"])) if "reference_compute_s" in row else ""),
          file=sys.stderr)
print(json.dumps({"workload": "kmc_rates_b200.synth.landscape_2d(side, T=0.3, seed=0), |A|=|B|=1", "rows": rows}))
