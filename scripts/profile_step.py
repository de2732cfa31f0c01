"""Minimal driver for ncu: create one handle, run `--warmup` + `--steps` numeric steps of a workload.
    python scripts/profile_step.py c2            # 10^5-minima landscape (bench.py's N=1 workload)
    python scripts/profile_step.py dense 8192    # complete graph
    python scripts/profile_step.py batch 256 512 # 512 complete 256-node networks
    python scripts/profile_step.py c4            # 2*10^4-node landscape, |A|=|B|=512 (phase 2)
    python scripts/profile_step.py c5            # the 4096 x 256-node temperature sweep of bench.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("size", nargs="?", type=int, default=0)
ap.add_argument("nbatch", nargs="?", type=int, default=0)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--gemm", type=int, default=0)
ap.add_argument("--leaf", type=int, default=0, help="nested-dissection leaf size (0 = the default, 64)")
a = ap.parse_args()
if a.workload == "c2":
    n, src, dst, k, A, B = synth.landscape_2d(a.size or 316, T=0.3, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B, gemm_path=a.gemm, leaf_size=a.leaf)
elif a.workload == "c4":
    n, src, dst, k, A, B = synth.landscape_groups(141, 512, 512, T=0.3, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B, gemm_path=a.gemm)
elif a.workload == "c5":
    import bench
    K = bench.c5_rate_matrices(torch, 0, a.nbatch or bench.C5_NB)
    h = Handle.from_dense(K, [0], [bench.C5_N - 1], gemm_path=a.gemm)
elif a.workload == "dist":
    # the multi-GPU dense path with a world of one rank (same launches per block as every rank runs at N > 1)
    from kmc_rates_b200 import _capi
    N = a.size or 8192
    K = torch.rand(N, N, dtype=torch.float64, device="cuda"); K.fill_diagonal_(0)
    h = Handle.from_dense(K, [0], [1], gemm_path=a.gemm)
    comm = _capi.Comm(0, 0, 1, lambda b: b)
    for _ in range(a.warmup):
        h.dist_phase_one(comm)
    ms = h.dist_phase_one(comm)
    print("workload dist n=%d: %.3f ms" % (N, ms))
    sys.exit(0)
elif a.workload == "dense":
    N = a.size or 8192
    K = torch.rand(N, N, dtype=torch.float64, device="cuda"); K.fill_diagonal_(0)
    h = Handle.from_dense(K, [0], [1], gemm_path=a.gemm)
else:
    N, nb = a.size or 256, a.nbatch or 512
    K = torch.rand(nb, N, N, dtype=torch.float64, device="cuda")
    h = Handle.from_dense(K, [0], [N - 1], gemm_path=a.gemm)
for _ in range(a.warmup):
    h.timed_steps(1)
ms = h.timed_steps(a.steps)
st = h.stats()
print("workload %s: %.3f ms/step, %d launches/step, schur flops %.3g, alg flops %.3g" % (a.workload, ms / a.steps, st["n_launches"], st["schur_flops"], st["alg_flops"]))
