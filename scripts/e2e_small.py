"""End-to-end cost of a SMALL network (the same-size leg of bench.py: 4 900 minima): where do the milliseconds go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle
side = int(sys.argv[1]) if len(sys.argv) > 1 else 70
n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0)
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(5):
    t0 = time.perf_counter(); h = Handle.from_coo(n, src, dst, k, A, B)
    t1 = time.perf_counter(); h.compute_rates(); t2 = time.perf_counter(); r = h.rates(); t3 = time.perf_counter(); h.close(); t4 = time.perf_counter()
    print("rep %d: create %.2f ms compute %.2f ms getters %.2f ms close %.2f ms total %.2f ms  (launches %d)" % (rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0), 0))
