"""Where does an end-to-end call spend its time? (host side; run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle
n, src, dst, k, A, B = synth.landscape_2d(316, T=0.3, seed=0)
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(4):
    t0 = time.perf_counter(); h = Handle.from_coo(n, src, dst, k, A, B); torch.cuda.synchronize()
    t1 = time.perf_counter(); h.compute_rates(); t2 = time.perf_counter(); r = h.rates(); t3 = time.perf_counter(); h.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print("rep %d: create %.1f ms (internal %.1f) compute %.1f ms getters %.2f ms close %.1f ms total %.1f ms" % (rep, 1e3*(t1-t0), 0.0, 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0)))
