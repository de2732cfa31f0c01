"""BASELINE config 4 at full size: 2*10^4-node landscape, |A| = |B| = 512; phase-2 per-a reductions on one GPU,
checked against the closed-form identity 1-P'_aa = 1/G_aa, tau'_a = (G tau_A)_a / G_aa with G = (I - P_AA)^-1
 evaluated in numpy from the reduced block."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle
side, g = int(sys.argv[1]) if len(sys.argv) > 1 else 141, int(sys.argv[2]) if len(sys.argv) > 2 else 512
n, src, dst, k, A, B = synth.landscape_groups(side, g, g, T=0.3, seed=0)
t0 = time.perf_counter(); h = Handle.from_coo(n, src, dst, k, A, B); t1 = time.perf_counter()
h.compute_rates(); t2 = time.perf_counter()
h.phase_one(); h.phase_two(); st = h.stats()
print("config4: n=%d |A|=|B|=%d create %.3f s compute %.3f s  phase1 %.2f ms phase2 %.2f ms launches %d" % (n, g, t1 - t0, t2 - t1, st["ms_phase1"], st["ms_phase2"], st["n_launches"]))
om, tau = h.final(); P, t = h.reduced(); r = h.rates()
for name, lo, hi in (("A", 0, g), ("B", g, 2 * g)):
    G = np.linalg.inv(np.eye(g) - P[lo:hi, lo:hi])
    om_cf = 1.0 / np.diag(G); tau_cf = (G @ t[lo:hi]) / np.diag(G)
    print("  group %s: max rel err vs closed form: 1-P'aa %.2e  tau' %.2e" % (name, np.max(np.abs(om[lo:hi] - om_cf) / om_cf), np.max(np.abs(tau[lo:hi] - tau_cf) / tau_cf)))
print("  kAB %.12g kBA %.12g kAB_SS %.12g kBA_SS %.12g" % tuple(r))
