"""Scratch diagnostics for the GPU box: ladder of cases, prints errors instead of asserting."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmc_rates_b200 import NGT, synth
from oracle.cpu_ngt import CpuNGT

def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

def run(name, make_gpu, make_cpu, committors=False):
    try:
        t0 = time.time(); g = make_gpu(); t1 = time.time()
        if committors: g.compute_rates_and_committors()
        else: g.compute_rates()
        t2 = time.time()
        got = np.array([g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()])
        c = make_cpu()
        if committors: c.compute_rates_and_committors()
        else: c.compute_rates()
        want = c.rates()
        msg = "%-34s rates err %.2e" % (name, rel(got, want))
        om, tau = g.get_final(); com, ctau = c.final()
        msg += "  final err %.2e %.2e" % (rel(om, com), rel(tau, ctau))
        if committors:
            q = g.get_committor_array(); cq = c.committors(); ok = ~np.isnan(cq)
            msg += "  q err %.2e" % float(np.max(np.abs(q[ok] - cq[ok])))
        st = g.stats()
        msg += "  create %.3fs compute %.3fs [p1 %.3f ms p2 %.3f ms launches %d fronts %d]" % (t1 - t0, t2 - t1, st["ms_phase1"], st["ms_phase2"], st["n_launches"], st["n_fronts"])
        print(msg, flush=True)
        if rel(got, want) > 1e-9: print("   got ", got, "\n   want", want)
    except Exception:
        print("%-34s EXCEPTION" % name); traceback.print_exc()

def dense_K(n, s, d, k):
    K = np.zeros((n, n)); K[s, d] = k; return K

for gp in (1, 0):
    tag = "dfma" if gp else "dmma"
    for n in (3, 10, 33, 40, 100, 130, 300):
        s, d, k = synth.complete_random(n, seed=n)
        A, B = [0], [n - 1]
        run("dense n=%d %s" % (n, tag), lambda: NGT.from_dense(dense_K(n, s, d, k), A, B, gemm_path=gp), lambda: CpuNGT(n, s, d, k, A, B), committors=(n <= 100))
        run("coo-complete n=%d %s" % (n, tag), lambda: NGT.from_coo(n, s, d, k, A, B, gemm_path=gp), lambda: CpuNGT(n, s, d, k, A, B))
    s, d, k = synth.complete_random(50, seed=1)
    run("dense n=50 A3 B4 %s" % tag, lambda: NGT.from_dense(dense_K(50, s, d, k), [0, 1, 2], [3, 4, 5, 6], gemm_path=gp), lambda: CpuNGT(50, s, d, k, [0, 1, 2], [3, 4, 5, 6]))
    s, d, k = synth.complete_random(120, seed=1)
    run("dense n=120 A40 B40 %s" % tag, lambda: NGT.from_dense(dense_K(120, s, d, k), list(range(40)), list(range(80, 120)), gemm_path=gp), lambda: CpuNGT(120, s, d, k, list(range(40)), list(range(80, 120))))
    for side, leaf in ((6, 8), (12, 16), (24, 16), (24, 64), (50, 64)):
        n, s, d, k, A, B = synth.landscape_2d(side, T=0.3, seed=side)
        run("landscape side=%d leaf=%d %s" % (side, leaf, tag), lambda: NGT.from_coo(n, s, d, k, A, B, gemm_path=gp, leaf_size=leaf), lambda: CpuNGT(n, s, d, k, A, B), committors=(side <= 24))
# bigger
n, s, d, k, A, B = synth.landscape_2d(100, T=0.3, seed=0)
run("landscape side=100", lambda: NGT.from_coo(n, s, d, k, A, B), lambda: CpuNGT(n, s, d, k, A, B, kind="reference" if os.path.exists("oracle/_ref/libngt_ref.so") else "port"))
t0 = time.time(); n, s, d, k, A, B = synth.landscape_2d(316, T=0.3, seed=0); print("gen 316: %.2fs" % (time.time() - t0))
try:
    t0 = time.time(); g = NGT.from_coo(n, s, d, k, A, B); t1 = time.time(); g.compute_rates(); t2 = time.time()
    print("landscape side=316: create %.3fs compute %.3fs kAB=%.15g kBA=%.15g" % (t1 - t0, t2 - t1, g.get_rate_AB(), g.get_rate_BA()), g.stats())
    for i in range(3):
        t0 = time.time(); g._h.phase_one(); t1 = time.time(); print("  phase_one again: %.4fs" % (t1 - t0), g.stats()["ms_phase1"], g.stats()["ms_ingest"])
    from oracle.py_linalg import TwoStateRatesCOO
    t0 = time.time(); la = TwoStateRatesCOO(n, s, d, k, A, B); la.compute_rates(); print("  linalg kAB=%.15g (%.2fs)" % (la.get_rate_AB(), time.time() - t0))
except Exception:
    traceback.print_exc()
import torch
for N in (2048, 4096, 8192):
    try:
        K = torch.rand(N, N, dtype=torch.float64, device="cuda"); K.fill_diagonal_(0)
        t0 = time.time(); g = NGT.from_dense(K, [0], [1]); g.compute_rates(); torch.cuda.synchronize(); t1 = time.time()
        st = g.stats()
        print("dense N=%d: total %.3fs p1 %.2f ms (%.2f TFLOP/s schur, flops %.3g) kAB=%.12g launches %d" % (N, t1 - t0, st["ms_phase1"], st["schur_flops"] / st["ms_phase1"] / 1e9, st["schur_flops"], g.get_rate_AB(), st["n_launches"]))
        if N <= 4096:
            Kc = K.cpu().numpy(); M = Kc - np.diag(Kc.sum(1)); idx = [i for i in range(N) if i != 1]
            t = np.linalg.solve(M[np.ix_(idx, idx)], -np.ones(N - 1)); print("   linear-solve kAB=%.12g" % (1.0 / t[0]))
    except Exception:
        traceback.print_exc()
