"""Scalability probe: landscape with side^2 minima (default 1000 -> 10^6 minima, ~10^7 transition states)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.perf_counter(); n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0); t1 = time.perf_counter()
host_sweeps = int(os.environ.get("HOST_SWEEPS", "0"))
h = Handle.from_coo(n, src, dst, k, A, B, host_sweeps=host_sweeps); h.close()          # cold: workspaces, device context
t1 = time.perf_counter(); h = Handle.from_coo(n, src, dst, k, A, B, host_sweeps=host_sweeps); t2 = time.perf_counter()
h.compute_rates(); t3 = time.perf_counter()
ms = h.timed_steps(3) / 3
st = h.stats()
print("n=%d rates=%d: generate %.2fs create %.2fs first compute %.3fs; step %.2f ms (%.3g nodes/s); fronts %d levels %d max front %d pool %.2f GB; kAB=%.12g kBA=%.12g" % (
    n, src.size, t1 - t0, t2 - t1, t3 - t2, ms, st["n_eliminated"] / (ms * 1e-3), st["n_fronts"], st["n_levels"], st["max_front"], st["pool_doubles"] * 8 / 1e9, h.rates()[0], h.rates()[1]))
if side <= 400:
    from oracle.py_linalg import TwoStateRatesCOO
    la = TwoStateRatesCOO(n, src, dst, k, A, B); la.compute_rates(); print("  linear solve kAB=%.12g" % la.get_rate_AB())
